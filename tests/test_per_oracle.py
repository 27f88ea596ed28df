"""CPU: the PER oracle vs fixtures produced by the reference's own SumSegmentTree / MinSegmentTree (oracle/make_golden.py)."""
import os

import numpy as np

from oracle.per_oracle import PerOracle
from tests.conftest import GOLDEN


def replay_case(z, ci, sampler):
    mem, alpha, beta, nadd, batch, seed, size, max_p = z[f'c{ci}_meta']
    sampler.add(int(nadd))
    sampler.update_priorities(z[f'c{ci}_upd_idx'], z[f'c{ci}_upd_p'])
    sampler.add(5)
    return sampler.sample(z[f'c{ci}_u'], float(beta))


def test_per_oracle_matches_reference_trees():
    z = np.load(os.path.join(GOLDEN, 'per_cases.npz'))
    n = len([k for k in z.files if k.endswith('_meta')])
    assert n == 3
    for ci in range(n):
        mem, alpha, beta, nadd, batch, seed, size, max_p = z[f'c{ci}_meta']
        o = PerOracle(int(mem), float(alpha))
        idxs, w = replay_case(z, ci, o)
        assert o.size == int(size) and o.max_priority == float(max_p)
        assert np.array_equal(idxs, z[f'c{ci}_idxs'])                  # integer work: bit exact
        assert np.array_equal(w, z[f'c{ci}_w'])                        # same float64 statements in the same order
        assert o.sum_tree.operate() == z[f'c{ci}_sum_root'][0] and o.min_tree.operate() == z[f'c{ci}_min_root'][0]
