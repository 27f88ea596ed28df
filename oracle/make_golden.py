"""Generate tests/golden/*.npz by running the REFERENCE's own importable modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

It imports, by path, the three leaf modules of the reference hot path
  /root/reference/scalerl/algorithms/impala/vtrace.py
  /root/reference/scalerl/algorithms/impala/loss_fn.py
  /root/reference/scalerl/algorithms/utils/atari_model.py
(the trainer module impala_atari.py itself cannot be imported: wrong import roots + gymnasium missing,
SURVEY.md §0) and drives them with the statements of ImpalaTrainer.learn (impala_atari.py:288-346).
Inputs are NOT stored: they are regenerated from seeds by oracle.impala_oracle.{init_params,
synthetic_batch} (numpy RNG), so the fixtures stay a few KB.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import impala_oracle as O  # noqa: E402

REF = '/root/reference/scalerl/algorithms'


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.set_num_threads(8)
    vtrace = _load('ref_vtrace', f'{REF}/impala/vtrace.py')
    loss_fn = _load('ref_loss_fn', f'{REF}/impala/loss_fn.py')
    atari_model = _load('ref_atari_model', f'{REF}/utils/atari_model.py')
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)

    # ---------------- V-trace known-answer vectors (vtrace.py:78-172) ----------------
    vt = {}
    cases = [(5, 3, 1.0, 1.0, 0), (20, 32, 1.0, 1.0, 1), (20, 7, None, None, 2), (33, 5, 2.0, 0.5, 3),
             (100, 4, 1.0, 1.0, 4), (1, 1, 1.0, 1.0, 5), (64, 2, 1.0, None, 6)]
    for i, (T, B, cr, cp, seed) in enumerate(cases):
        rng = np.random.RandomState(seed)
        log_rhos = (rng.randn(T, B) * 0.7).astype(np.float32)
        discounts = ((rng.rand(T, B) > 0.1) * 0.99).astype(np.float32)
        rewards = rng.randn(T, B).astype(np.float32)
        values = rng.randn(T, B).astype(np.float32)
        boot = rng.randn(B).astype(np.float32)
        r = vtrace.from_importance_weights(torch.from_numpy(log_rhos), torch.from_numpy(discounts),
                                           torch.from_numpy(rewards), torch.from_numpy(values),
                                           torch.from_numpy(boot), clip_rho_threshold=cr, clip_pg_rho_threshold=cp)
        vt[f'c{i}_meta'] = np.array([T, B, -1 if cr is None else cr, -1 if cp is None else cp, seed], dtype=np.float64)
        for k, v in dict(log_rhos=log_rhos, discounts=discounts, rewards=rewards, values=values, boot=boot,
                         vs=r.vs.numpy(), pg=r.pg_advantages.numpy()).items():
            vt[f'c{i}_{k}'] = v
    np.savez_compressed(os.path.join(out_dir, 'vtrace_cases.npz'), **vt)

    # ---------------- whole learn() step through the reference modules ----------------
    step_cases = [dict(name='t5b4a6', T=5, B=4, A=6, seed=0, reward_clipping='abs_one', steps=2),
                  dict(name='t3b5a4', T=3, B=5, A=4, seed=1, reward_clipping='none', steps=1)]
    for c in step_cases:
        T, B, A = c['T'], c['B'], c['A']
        hp = dict(O.DEFAULT_HP, reward_clipping=c['reward_clipping'])
        params = O.init_params(A, seed=c['seed'])
        model = atari_model.AtariNet((4, 84, 84), A, use_lstm=False)
        model.load_state_dict(params)
        model.train()
        opt = torch.optim.RMSprop(model.parameters(), lr=hp['learning_rate'], momentum=hp['momentum'],
                                  eps=hp['epsilon'], alpha=hp['alpha'])            # impala_atari.py:99-105
        g = {}
        for step in range(c['steps']):
            batch = O.synthetic_batch(T, B, A, seed=c['seed'] * 10 + step)
            torch.manual_seed(0)
            learner_outputs, _ = model(batch, ())                                  # :289
            bootstrap_value = learner_outputs['baseline'][-1]                      # :293
            b1 = {k: t[1:] for k, t in batch.items()}                              # :296
            lo = {k: t[:-1] for k, t in learner_outputs.items()}                   # :297-300
            rewards = b1['reward']
            clipped = torch.clamp(rewards, -1, 1) if hp['reward_clipping'] == 'abs_one' else rewards
            discounts = (~b1['done']).float() * hp['discounting']                  # :308
            vr = vtrace.from_logits(behavior_policy_logits=b1['policy_logits'],
                                    target_policy_logits=lo['policy_logits'], actions=b1['action'],
                                    discounts=discounts, rewards=clipped, values=lo['baseline'],
                                    bootstrap_value=bootstrap_value)               # :310-318
            pg_loss = loss_fn.compute_policy_gradient_loss(lo['policy_logits'], b1['action'], vr.pg_advantages)
            baseline_loss = hp['baseline_cost'] * loss_fn.compute_baseline_loss(vr.vs - lo['baseline'])
            entropy_loss = hp['entropy_cost'] * loss_fn.compute_entropy_loss(lo['policy_logits'])
            total_loss = pg_loss + baseline_loss + entropy_loss                    # :330
            opt.zero_grad()
            total_loss.backward()                                                  # :343
            grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            gn = torch.nn.utils.clip_grad_norm_(model.parameters(), hp['max_grad_norm'])  # :344
            opt.step()                                                             # :346
            s = f's{step}_'
            g[s + 'policy_logits'] = learner_outputs['policy_logits'].detach().numpy()
            g[s + 'baseline'] = learner_outputs['baseline'].detach().numpy()
            g[s + 'vs'] = vr.vs.numpy()
            g[s + 'pg_advantages'] = vr.pg_advantages.numpy()
            g[s + 'log_rhos'] = vr.log_rhos.detach().numpy()
            g[s + 'losses'] = np.array([pg_loss.item(), baseline_loss.item(), entropy_loss.item(), total_loss.item()])
            g[s + 'grad_norm'] = np.array([float(gn)])
            for k, v in grads.items():
                flat = v.reshape(-1).double()
                g[s + 'gradnorm_' + k] = np.array([float(flat.norm())])
                g[s + 'gradsum_' + k] = np.array([float(flat.sum())])
                g[s + 'gradhead_' + k] = v.reshape(-1)[:64].numpy().copy()
                # a strided sample across the whole tensor
                idx = torch.linspace(0, flat.numel() - 1, steps=min(257, flat.numel())).long()
                g[s + 'gradsamp_' + k] = v.reshape(-1)[idx].numpy().copy()
            for k, p in model.named_parameters():
                flat = p.detach().reshape(-1)
                idx = torch.linspace(0, flat.numel() - 1, steps=min(257, flat.numel())).long()
                g[s + 'param_' + k] = flat[idx].numpy().copy()
                g[s + 'paramsum_' + k] = np.array([float(flat.double().sum())])
        g['meta'] = np.array([T, B, A, c['seed'], c['steps'], 1 if c['reward_clipping'] == 'abs_one' else 0])
        np.savez_compressed(os.path.join(out_dir, f"learn_{c['name']}.npz"), **g)
        print('wrote', c['name'], 'total_loss', float(total_loss))

    # ---------------- LSTM core through the reference AtariNet(use_lstm=True) (atari_model.py:52-55,109-120) ----------------
    T, B, A = 4, 3, 6
    params = O.init_params(A, seed=2)
    lp = O.init_lstm_params(A, seed=2)
    model = atari_model.AtariNet((4, 84, 84), A, use_lstm=True)
    model.load_state_dict({**params, **lp})
    model.train()
    batch = O.synthetic_batch(T, B, A, seed=42, done_p=0.25)
    rng = np.random.RandomState(5)
    state = (torch.from_numpy(rng.randn(2, B, 513 + A).astype(np.float32) * 0.3), torch.from_numpy(rng.randn(2, B, 513 + A).astype(np.float32) * 0.3))
    torch.manual_seed(0)
    out, new_state = model(batch, state)
    baseline = out['baseline']
    tl, tv = out['policy_logits'][:-1], baseline[:-1]
    b1 = {k: t[1:] for k, t in batch.items()}
    discounts = (~b1['done']).float() * 0.99
    vr = vtrace.from_logits(behavior_policy_logits=b1['policy_logits'], target_policy_logits=tl, actions=b1['action'], discounts=discounts,
                            rewards=torch.clamp(b1['reward'], -1, 1), values=tv, bootstrap_value=baseline[-1])
    total = (loss_fn.compute_policy_gradient_loss(tl, b1['action'], vr.pg_advantages) + 0.5 * loss_fn.compute_baseline_loss(vr.vs - tv) +
             0.0006 * loss_fn.compute_entropy_loss(tl))
    model.zero_grad()
    total.backward()
    g = {'policy_logits': out['policy_logits'].detach().numpy(), 'baseline': baseline.detach().numpy(), 'vs': vr.vs.numpy(),
         'h_out': new_state[0].detach().numpy(), 'c_out': new_state[1].detach().numpy(), 'total_loss': np.array([total.item()]),
         'state_h': state[0].numpy(), 'state_c': state[1].numpy(), 'meta': np.array([T, B, A, 2, 42])}
    for k, p_ in model.named_parameters():
        flat = p_.grad.detach().reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, steps=min(257, flat.numel())).long()
        g['gradsamp_' + k] = flat[idx].numpy().copy()
        g['gradnorm_' + k] = np.array([float(flat.double().norm())])
    np.savez_compressed(os.path.join(out_dir, 'lstm_t4b3a6.npz'), **g)
    print('wrote lstm_t4b3a6 total_loss', total.item())

    # ---------------- prioritized replay: the reference's own tree classes driven by the statements of
    # PrioritizedReplayBuffer._add / update_priorities / _sample_proprtional / _calculate_weight (replay_buffer.py:318-381);
    # `retrieve` (missing upstream) -> find_prefixsum_idx ----------------
    seg = _load('ref_segment_tree', '/root/reference/scalerl/data/segment_tree.py')
    per = {}
    for ci, (mem, alpha, beta, nadd, batch, seed) in enumerate([(100, 0.6, 0.4, 100, 32, 0), (1000, 0.7, 0.5, 700, 64, 1), (64, 0.6, 1.0, 200, 16, 2)]):
        cap = 1
        while cap < mem:
            cap *= 2
        st, mt = seg.SumSegmentTree(cap), seg.MinSegmentTree(cap)
        rng = np.random.RandomState(seed)
        max_p, ptr, size = 1.0, 0, 0
        for _ in range(nadd):
            st[ptr] = max_p ** alpha
            mt[ptr] = max_p ** alpha
            ptr = (ptr + 1) % mem
            size = min(size + 1, mem)
        upd_idx = rng.randint(0, size, size=3 * batch)
        upd_p = rng.rand(3 * batch) * 5 + 1e-3
        for i, pr in zip(upd_idx, upd_p):
            st[int(i)] = float(pr) ** alpha
            mt[int(i)] = float(pr) ** alpha
            max_p = max(max_p, float(pr))
        for _ in range(5):                                   # adds after updates use the new max priority
            st[ptr] = max_p ** alpha
            mt[ptr] = max_p ** alpha
            ptr = (ptr + 1) % mem
            size = min(size + 1, mem)
        u = rng.rand(batch)
        p_total = st.sum(0, size - 1)
        segment = p_total / batch
        idxs = []
        for i in range(batch):
            a, b = segment * i, segment * (i + 1)
            idxs.append(st.find_prefixsum_idx(a + (b - a) * float(u[i])))
        p_min = mt.min() / st.sum()
        max_w = (p_min * size) ** (-beta)
        w = [((st[i] / st.sum()) * size) ** (-beta) / max_w for i in idxs]
        per[f'c{ci}_meta'] = np.array([mem, alpha, beta, nadd, batch, seed, size, max_p], dtype=np.float64)
        per[f'c{ci}_upd_idx'], per[f'c{ci}_upd_p'], per[f'c{ci}_u'] = upd_idx.astype(np.int64), upd_p, u
        per[f'c{ci}_idxs'], per[f'c{ci}_w'] = np.array(idxs, dtype=np.int64), np.array(w, dtype=np.float64)
        per[f'c{ci}_sum_root'], per[f'c{ci}_min_root'] = np.array([st.sum()]), np.array([mt.min()])
    np.savez_compressed(os.path.join(out_dir, 'per_cases.npz'), **per)
    print('wrote per_cases')


if __name__ == '__main__':
    main()
