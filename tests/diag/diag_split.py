"""Where does the fp32-accurate (split) mode lose accuracy?  Every intermediate of one learner step (hi + lo tensors) against
fp32 torch autograd on the CPU.  Run on the GPU box:  python tests/diag/diag_split.py [T B A]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import impala_oracle as O          # noqa: E402
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams   # noqa: E402
from tests.helpers import rel_l2               # noqa: E402


def main():
    T, B, A = [int(x) for x in sys.argv[1:4]] if len(sys.argv) >= 4 else (5, 4, 6)
    prec = sys.argv[4] if len(sys.argv) > 4 else 'fp32_split'
    params = O.init_params(A, seed=1)
    batch = O.synthetic_batch(T, B, A, seed=10)
    L = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A, precision=prec, learning_rate=0.0), init_state_dict=params,
                          process_group=False, use_graph=False)
    L.learn({k: v.cuda() for k, v in batch.items()})
    NF, NB = (T + 1) * B, T * B
    # ---- CPU fp32 autograd with retained intermediates
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x = batch['obs'].reshape(NF, 4, 84, 84).float() / 255.0
    a1 = F.relu(F.conv2d(x, ps['conv1.weight'], ps['conv1.bias'], stride=4)); a1.retain_grad()
    a2 = F.relu(F.conv2d(a1, ps['conv2.weight'], ps['conv2.bias'], stride=2)); a2.retain_grad()
    a3 = F.relu(F.conv2d(a2, ps['conv3.weight'], ps['conv3.bias'], stride=1)); a3.retain_grad()
    h = F.relu(F.linear(a3.reshape(NF, -1), ps['fc.weight'], ps['fc.bias'])); h.retain_grad()
    core = torch.cat([h, torch.clamp(batch['reward'], -1, 1).reshape(NF, 1), F.one_hot(batch['action'].reshape(NF), A).float()], -1)
    logits = F.linear(core, ps['policy.weight'], ps['policy.bias']).view(T + 1, B, A)
    base = F.linear(core, ps['baseline.weight'], ps['baseline.bias']).view(T + 1, B)
    rewards = torch.clamp(batch['reward'][1:], -1, 1)
    disc = (~batch['done'][1:]).float() * 0.99
    with torch.no_grad():
        vs, pg, *_ = O.vtrace_from_logits(batch['policy_logits'][1:], logits[:-1], batch['action'][1:], disc, rewards, base[:-1], base[-1])
    l1, l2, l3 = O.impala_losses(logits[:-1], batch['action'][1:], base[:-1], vs, pg, 0.5, 0.0006)
    (l1 + l2 + l3).backward()

    def dbg(name):
        t = L.debug_buffer(name).float()
        if prec == 'fp32_split':
            t = t + L.debug_buffer(name + '_lo').float()
        return t.cpu()
    rep = {}
    a1g = dbg('a1').view(2, NF, 10, 10, 2, 32).permute(1, 5, 2, 0, 3, 4).reshape(NF, 32, 20, 20)
    rep['a1'] = rel_l2(a1g, a1.detach())
    rep['a2'] = rel_l2(dbg('a2').view(NF, 9, 9, 64).permute(0, 3, 1, 2), a2.detach())
    rep['a3'] = rel_l2(dbg('a3').view(NF, 7, 7, 64).permute(0, 3, 1, 2), a3.detach())
    rep['h'] = rel_l2(L.debug_buffer('h').view(NF, 512).cpu(), h.detach())
    rep['dh'] = rel_l2(dbg('dh').view(NB, 512), (h.grad * (h > 0))[:NB])
    # gradients of the PRE-activations: d a_k * (a_k > 0)
    g3 = (a3.grad * (a3 > 0))[:NB]
    rep['da3'] = rel_l2(dbg('da3').view(NB, 9, 9, 64)[:, :7, :7].permute(0, 3, 1, 2), g3)
    g2 = (a2.grad * (a2 > 0))[:NB]
    rep['da2'] = rel_l2(dbg('da2').view(NB, 10, 10, 64)[:, :9, :9].permute(0, 3, 1, 2), g2)
    g1 = (a1.grad * (a1 > 0))[:NB]
    rep['da1'] = rel_l2(dbg('da1').view(NB, 21, 21, 32)[:, :20, :20, :].permute(0, 3, 1, 2), g1)
    for k in O.PARAM_ORDER:
        rep['grad ' + k] = rel_l2(L.grads[k].cpu(), ps[k].grad)
    for k, v in rep.items():
        print(f'{k:24s} {v:.3e}')


if __name__ == '__main__':
    main()
