"""GPU parity of the prioritized-replay sampler (configs[3]) against the reference-pinned oracle: integer results (sampled
indices) bit-exact, float64 tree contents exact given identical leaves, IS weights within 1e-12."""
import os

import numpy as np
import pytest
import torch

from oracle.per_oracle import PerOracle
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _run_case(z, ci):
    from scalerl_b200.data.per_sampler import GpuPrioritizedSampler
    mem, alpha, beta, nadd, batch, seed, size, max_p = z[f'c{ci}_meta']
    s = GpuPrioritizedSampler(int(mem), float(alpha))
    s.add(int(nadd))
    s.update_priorities(torch.from_numpy(z[f'c{ci}_upd_idx']), torch.from_numpy(z[f'c{ci}_upd_p']))
    s.add(5)
    idxs, w32 = s.sample(int(batch), float(beta), uniforms=torch.from_numpy(z[f'c{ci}_u']))
    return s, idxs.cpu().numpy(), s._w64.cpu().numpy(), w32.cpu().numpy()


def test_per_matches_reference_goldens():
    z = np.load(os.path.join(GOLDEN, 'per_cases.npz'))
    for ci in range(3):
        mem, alpha, beta, nadd, batch, seed, size, max_p = z[f'c{ci}_meta']
        s, idxs, w64, w32 = _run_case(z, ci)
        assert len(s) == int(size)
        st, mt, mp = s.trees()
        assert mp == float(max_p)
        assert abs(st[1].item() - z[f'c{ci}_sum_root'][0]) <= 1e-12 * z[f'c{ci}_sum_root'][0]
        assert abs(mt[1].item() - z[f'c{ci}_min_root'][0]) <= 1e-14 * z[f'c{ci}_min_root'][0]
        assert np.array_equal(idxs, z[f'c{ci}_idxs']), (ci, int((idxs != z[f'c{ci}_idxs']).sum()))
        assert np.allclose(w64, z[f'c{ci}_w'], rtol=1e-12, atol=0)
        assert np.allclose(w32, z[f'c{ci}_w'], rtol=1e-6)


@pytest.mark.parametrize('mem,batch,rounds', [(1 << 16, 512, 4), (50000, 64, 3), (2, 4, 1)])
def test_per_random_workload_vs_oracle(mem, batch, rounds):
    """larger trees, duplicate indices inside one update (last write wins), ring wrap-around; the oracle is rebuilt from the
    GPU's own leaf values so that index equality is exact (device pow() may differ from libm's by an ulp)"""
    from scalerl_b200.data.per_sampler import GpuPrioritizedSampler
    rng = np.random.RandomState(mem % 1000)
    alpha, beta = 0.6, 0.4
    s = GpuPrioritizedSampler(mem, alpha)
    o = PerOracle(mem, alpha)
    s.add(mem + mem // 3)
    o.add(mem + mem // 3)
    for r in range(rounds):
        n = min(3 * batch, 2000)
        ui = rng.randint(0, len(s), size=n)
        if n > 10:
            ui[5] = ui[3]                                        # duplicates
            ui[n - 1] = ui[0]
        up = rng.rand(n) * 3 + 1e-3
        s.update_priorities(torch.from_numpy(ui), torch.from_numpy(up))
        o.update_priorities(ui, up)
        st, mt, mp = s.trees()
        assert mp == o.max_priority
        cap = s.capacity
        leaves = st[cap:cap + mem].cpu().numpy()
        ref_leaves = o.sum_tree.tree[cap:cap + mem]
        assert np.allclose(leaves, ref_leaves, rtol=1e-14, atol=0)
        o.set_leaves(np.nonzero(leaves != ref_leaves)[0], leaves[leaves != ref_leaves])     # adopt the GPU's rounding of p**alpha
        assert np.array_equal(st.cpu().numpy()[1:], o.sum_tree.tree[1:])                    # every internal node identical
        assert np.array_equal(mt.cpu().numpy()[1:], o.min_tree.tree[1:])
        u = rng.rand(batch)
        idxs, w32 = s.sample(batch, beta, uniforms=torch.from_numpy(u))
        ri, rw = o.sample(u, beta)
        assert np.array_equal(idxs.cpu().numpy(), ri)
        assert np.allclose(s._w64.cpu().numpy(), rw, rtol=1e-12, atol=0)
    # stratification property at full size: one sample per equal-mass segment -> indices are non-decreasing
    idxs, _ = s.sample(batch, beta)
    d = np.diff(idxs.cpu().numpy())
    assert (d >= 0).all()


def test_per_argument_errors():
    from scalerl_b200.data.per_sampler import GpuPrioritizedSampler
    s = GpuPrioritizedSampler(16)
    with pytest.raises(ValueError):
        s.sample(4)                       # empty buffer
    with pytest.raises(ValueError):
        GpuPrioritizedSampler(1)
