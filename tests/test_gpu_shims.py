"""The differentiable drop-ins of scalerl/algorithms/impala/{vtrace,loss_fn}.py: the reference's own learn statements
(impala_atari.py:289-346) run with ONLY the imports swapped, and total_loss.backward() yields the oracle's gradients."""
import pytest
import torch
import torch.nn.functional as F

from oracle import impala_oracle as O
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _inputs(T, B, A, seed):
    g = torch.Generator().manual_seed(seed)
    batch = O.synthetic_batch(T, B, A, seed=seed, done_p=0.1)
    lo = dict(policy_logits=torch.randn(T + 1, B, A, generator=g), baseline=torch.randn(T + 1, B, generator=g))
    return batch, lo


@pytest.mark.parametrize('T,B,A,clip', [(5, 4, 6, 'abs_one'), (20, 32, 6, 'abs_one'), (7, 3, 18, 'none'), (1, 1, 2, 'abs_one')])
def test_reference_learn_statements_with_the_drop_ins(T, B, A, clip):
    from scalerl_b200.algorithms.impala.vtrace import from_logits
    from scalerl_b200.algorithms.impala.loss_fn import compute_baseline_loss, compute_entropy_loss, compute_policy_gradient_loss
    baseline_cost, entropy_cost, discounting = 0.5, 0.0006, 0.99
    batch_cpu, lo_cpu = _inputs(T, B, A, seed=T * 100 + B)
    batch = {k: v.cuda() for k, v in batch_cpu.items()}
    learner_outputs = {k: v.cuda().requires_grad_(True) for k, v in lo_cpu.items()}
    full = dict(learner_outputs)
    # ---- impala_atari.py:293-330, verbatim in structure ------------------------------------------------------------
    bootstrap_value = learner_outputs['baseline'][-1]
    batch = {key: tensor[1:] for key, tensor in batch.items()}
    learner_outputs = {key: tensor[:-1] for key, tensor in learner_outputs.items()}
    rewards = batch['reward']
    clipped_rewards = torch.clamp(rewards, -1, 1) if clip == 'abs_one' else rewards
    discounts = (~batch['done']).float() * discounting
    vtrace_returns = from_logits(behavior_policy_logits=batch['policy_logits'], target_policy_logits=learner_outputs['policy_logits'],
                                 actions=batch['action'], discounts=discounts, rewards=clipped_rewards,
                                 values=learner_outputs['baseline'], bootstrap_value=bootstrap_value)
    pg_loss = compute_policy_gradient_loss(learner_outputs['policy_logits'], batch['action'], vtrace_returns.pg_advantages)
    baseline_loss = baseline_cost * compute_baseline_loss(vtrace_returns.vs - learner_outputs['baseline'])
    entropy_loss = entropy_cost * compute_entropy_loss(learner_outputs['policy_logits'])
    total_loss = pg_loss + baseline_loss + entropy_loss
    total_loss.backward()
    # ---- the oracle on the same inputs --------------------------------------------------------------------------------
    tl, tb = lo_cpu['policy_logits'], lo_cpu['baseline']
    rw = batch_cpu['reward'][1:]
    rw = torch.clamp(rw, -1, 1) if clip == 'abs_one' else rw
    disc = (~batch_cpu['done'][1:]).float() * discounting
    vs, pg, lr, balp, talp = O.vtrace_from_logits(batch_cpu['policy_logits'][1:], tl[:-1], batch_cpu['action'][1:], disc, rw, tb[:-1], tb[-1])
    l_pg, l_bl, l_ent = O.impala_losses(tl[:-1], batch_cpu['action'][1:], tb[:-1], vs, pg, baseline_cost, entropy_cost)
    dl, dv = O.head_grads(tl[:-1], batch_cpu['action'][1:], tb[:-1], vs, pg, baseline_cost, entropy_cost)
    assert not vtrace_returns.vs.requires_grad and not vtrace_returns.pg_advantages.requires_grad       # vtrace.py:78
    assert vtrace_returns.log_rhos.requires_grad and vtrace_returns.target_action_log_probs.requires_grad
    assert_close(vtrace_returns.vs, vs, 1e-4, 'vs'); assert_close(vtrace_returns.pg_advantages, pg, 1e-4, 'pg_adv')
    assert_close(vtrace_returns.log_rhos, lr, 1e-4, 'log_rhos')
    for got, want, nm in ((pg_loss, l_pg, 'pg'), (baseline_loss, l_bl, 'baseline'), (entropy_loss, l_ent, 'entropy')):
        assert abs(float(got) - float(want)) <= 1e-4 * max(1.0, abs(float(want))), nm
    g_logits, g_base = full['policy_logits'].grad.cpu(), full['baseline'].grad.cpu()
    assert_close(g_logits[:-1], dl, 1e-4, 'd total / d logits')
    assert float(g_logits[-1].abs().max()) == 0.0                              # row T is dropped by [:-1]
    # d total / d baseline: -baseline_cost * (vs - V) for rows < T; the bootstrap row gets no gradient (vs is detached)
    assert_close(g_base[:-1], dv, 1e-4, 'd total / d baseline')
    assert float(g_base[-1].abs().max()) == 0.0


def test_each_drop_in_against_torch_autograd():
    from scalerl_b200.algorithms.impala import loss_fn as LF, vtrace as VT
    g = torch.Generator().manual_seed(5)
    T, B, A = 6, 5, 7
    logits = torch.randn(T, B, A, generator=g)
    actions = torch.randint(0, A, (T, B), generator=g)
    adv = torch.randn(T, B, generator=g)
    up = torch.randn(T, B, generator=g)

    def both(fn_mine, fn_ref, *extra):
        a = logits.clone().cuda().requires_grad_(True)
        b = logits.clone().requires_grad_(True)
        ym = fn_mine(a, *[e.cuda() for e in extra])
        yr = fn_ref(b, *extra)
        if ym.dim():
            ym.backward(up.cuda()); yr.backward(up)
        else:
            (3.0 * ym).backward(); (3.0 * yr).backward()
        assert_close(ym.detach(), yr.detach(), 1e-5, 'value')
        assert_close(a.grad, b.grad, 1e-5, 'grad')

    both(VT.action_log_probs, lambda l, a_: -F.nll_loss(F.log_softmax(torch.flatten(l, 0, -2), -1), torch.flatten(a_), reduction='none').view_as(a_), actions)
    both(LF.compute_entropy_loss, lambda l: torch.sum(F.softmax(l, -1) * F.log_softmax(l, -1)))
    both(LF.compute_policy_gradient_loss,
         lambda l, a_, ad: torch.sum(F.nll_loss(F.log_softmax(torch.flatten(l, 0, 1), -1), torch.flatten(a_, 0, 1), reduction='none').view_as(ad) * ad.detach()),
         actions, adv)
    x = adv.clone().cuda().requires_grad_(True)
    y = LF.compute_baseline_loss(x)
    (2.0 * y).backward()
    assert abs(float(y) - float(0.5 * (adv ** 2).sum())) < 1e-4
    assert_close(x.grad, 2.0 * adv, 1e-6, 'baseline grad')
