"""CPU: the oracle restatement vs fixtures produced by the reference's own modules
(oracle/make_golden.py; the reference has no tests/goldens of its own -- SURVEY.md §8c)."""
import os

import numpy as np
import pytest
import torch

from oracle import impala_oracle as O
from tests.conftest import GOLDEN
from tests.helpers import assert_close, rel_l2, strided_sample


def _vt_cases():
    z = np.load(os.path.join(GOLDEN, 'vtrace_cases.npz'))
    n = len([k for k in z.files if k.endswith('_meta')])
    return z, n


def test_vtrace_matches_reference_goldens():
    z, n = _vt_cases()
    assert n >= 7
    for i in range(n):
        T, B, cr, cp, _ = z[f'c{i}_meta']
        cr = None if cr < 0 else float(cr)
        cp = None if cp < 0 else float(cp)
        args = [torch.from_numpy(z[f'c{i}_{k}']) for k in ('log_rhos', 'discounts', 'rewards', 'values', 'boot')]
        vs, pg = O.vtrace_from_importance_weights(*args, cr, cp)
        # same torch ops, same order -> bit exact
        assert np.array_equal(vs.numpy(), z[f'c{i}_vs'])
        assert np.array_equal(pg.numpy(), z[f'c{i}_pg'])
        # the independent fp64 witness agrees to fp32 rounding
        vs64, pg64 = O.vtrace_from_importance_weights_np64(*[a.numpy() for a in args], cr, cp)
        assert_close(vs64, z[f'c{i}_vs'], 2e-6, f'vs64 c{i}')
        assert_close(pg64, z[f'c{i}_pg'], 2e-6, f'pg64 c{i}')


def test_vtrace_known_answers():
    # all-done (discount 0), log_rhos=0, r=1, V=0 => vs = pg_adv = 1
    T, B = 6, 3
    z = torch.zeros(T, B)
    vs, pg = O.vtrace_from_importance_weights(z, z, torch.ones(T, B), z, torch.zeros(B))
    assert torch.equal(vs, torch.ones(T, B)) and torch.equal(pg, torch.ones(T, B))
    # on-policy, no clipping active, gamma=1: vs_t = sum of future rewards + bootstrap
    r = torch.arange(T * B, dtype=torch.float32).view(T, B)
    vs, _ = O.vtrace_from_importance_weights(z, torch.ones(T, B), r, z, torch.full((B,), 2.0))
    expect = torch.flip(torch.cumsum(torch.flip(r, [0]), 0), [0]) + 2.0
    assert torch.allclose(vs, expect)


@pytest.mark.parametrize('name', ['t5b4a6', 't3b5a4'])
@pytest.mark.parametrize('use_autograd', [True, False])
def test_learn_step_matches_reference_goldens(name, use_autograd):
    g = np.load(os.path.join(GOLDEN, f'learn_{name}.npz'))
    T, B, A, seed, steps, clip = [int(v) for v in g['meta']]
    hp = dict(reward_clipping='abs_one' if clip else 'none')
    torch.set_num_threads(8)
    params = O.init_params(A, seed=seed)
    opt = O.new_opt_state(params)
    for step in range(steps):
        batch = O.synthetic_batch(T, B, A, seed=seed * 10 + step)
        out = O.learn_step(params, opt, batch, hp, use_autograd=use_autograd)
        s = f's{step}_'
        assert_close(out['policy_logits'], g[s + 'policy_logits'], 2e-6, 'logits')
        assert_close(out['baseline'], g[s + 'baseline'], 2e-6, 'baseline')
        assert_close(out['vs'], g[s + 'vs'], 5e-6, 'vs')
        assert_close(out['pg_advantages'], g[s + 'pg_advantages'], 5e-6, 'pg_adv')
        losses = np.array([out['pg_loss'], out['baseline_loss'], out['entropy_loss'], out['total_loss']])
        assert_close(losses, g[s + 'losses'], 1e-5, 'losses')
        assert abs(out['grad_norm'] - g[s + 'grad_norm'][0]) <= 1e-4 * g[s + 'grad_norm'][0]
        for k in O.PARAM_ORDER:
            flat = out['grads'][k].reshape(-1)
            scale = float(g[s + 'gradnorm_' + k][0]) / np.sqrt(flat.numel()) + 1e-12
            d = np.abs(strided_sample(flat).numpy() - g[s + 'gradsamp_' + k]).max()
            assert d <= 2e-4 * max(scale, np.abs(g[s + 'gradsamp_' + k]).max()), (k, d, scale)
            assert abs(float(flat.double().norm()) - g[s + 'gradnorm_' + k][0]) <= 1e-4 * g[s + 'gradnorm_' + k][0] + 1e-9
            p = params[k].reshape(-1)
            assert_close(strided_sample(p), g[s + 'param_' + k], 2e-6, 'param ' + k)


def test_manual_backward_equals_autograd():
    T, B, A = 4, 3, 6
    params = O.init_params(A, seed=3)
    batch = O.synthetic_batch(T, B, A, seed=7, done_p=0.2)
    a = O.learn_step(dict(params), None, batch, update=False, use_autograd=True)
    m = O.learn_step(dict(params), None, batch, update=False, use_autograd=False)
    for k in O.PARAM_ORDER:
        assert_close(m['grads'][k], a['grads'][k], 2e-5, k)


def test_bf16_emulation_is_close_to_fp32():
    T, B, A = 4, 3, 6
    params = O.init_params(A, seed=3)
    batch = O.synthetic_batch(T, B, A, seed=7)
    a = O.learn_step(dict(params), None, batch, update=False)
    e = O.learn_step(dict(params), None, batch, update=False, emulate_bf16=True)
    assert_close(e['policy_logits'], a['policy_logits'], 2e-2, 'logits bf16')
    for k in O.PARAM_ORDER:
        assert rel_l2(e['grads'][k], a['grads'][k]) < 0.12, k  # ReLU-mask flips: rel-L2 ~ sqrt(flip fraction)


def test_adam_step_matches_torch():
    params = {k: v.clone() for k, v in O.init_params(4, seed=1).items()}
    tp = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
    topt = torch.optim.Adam(tp.values(), lr=1e-3)
    st = O.new_opt_state(params, 'adam')
    rng = torch.Generator().manual_seed(0)
    for _ in range(3):
        grads = {k: torch.randn(v.shape, generator=rng) for k, v in params.items()}
        for k in tp:
            tp[k].grad = grads[k].clone()
        topt.step()
        st['step'] += 1
        O.adam_step(params, grads, st['exp_avg'], st['exp_avg_sq'], st['step'], 1e-3)
    for k in params:
        assert_close(params[k], tp[k].detach(), 1e-6, k)


def test_lstm_oracle_matches_reference_goldens():
    """use_lstm=True branch (atari_model.py:109-120): the oracle's step-wise LSTM with done-resets vs the reference AtariNet"""
    g = np.load(os.path.join(GOLDEN, 'lstm_t4b3a6.npz'))
    T, B, A, seed, bseed = [int(v) for v in g['meta']]
    params, lp = O.init_params(A, seed=seed), O.init_lstm_params(A, seed=seed)
    batch = O.synthetic_batch(T, B, A, seed=bseed, done_p=0.25)
    state = (torch.from_numpy(g['state_h']), torch.from_numpy(g['state_c']))
    with torch.no_grad():
        lg, bs, ns = O.atari_forward_lstm(params, lp, batch['obs'], batch['reward'], batch['action'], batch['done'], state)
    assert_close(lg, g['policy_logits'], 5e-6, 'logits')
    assert_close(bs, g['baseline'], 5e-6, 'baseline')
    assert_close(ns[0], g['h_out'], 5e-6, 'h')
    assert_close(ns[1], g['c_out'], 5e-6, 'c')
    out = O.learn_step_lstm(params, lp, batch, state)
    assert abs(out['total_loss'] - g['total_loss'][0]) <= 1e-4 * abs(g['total_loss'][0])
    assert_close(out['vs'], g['vs'], 1e-5, 'vs')
    for k, v in {**out['grads'], **out['lstm_grads']}.items():
        samp = strided_sample(v.reshape(-1)).numpy()
        scale = max(np.abs(g['gradsamp_' + k]).max(), float(g['gradnorm_' + k][0]) / np.sqrt(v.numel()))
        assert np.abs(samp - g['gradsamp_' + k]).max() <= 3e-4 * scale, k
