"""oracle/_ref (the reference's own vtrace / loss_fn / atari_model modules, built by oracle/make_ref.py) under the restated learn()
statements == the oracle's learn_step, whole step: losses, grad norm, post-step weights -- and it is what bench.py's reference arm runs."""
import pytest
import torch

from oracle import impala_oracle as O
from oracle import ref_learner as R

pytestmark = pytest.mark.skipif(not R.available(), reason='oracle/_ref not built (python oracle/make_ref.py, build container)')


@pytest.mark.parametrize('T,B,A,clip', [(5, 4, 6, 'abs_one'), (3, 7, 4, 'none')])
def test_reference_learner_step_equals_oracle(T, B, A, clip):
    torch.set_num_threads(4)
    params = O.init_params(A, seed=2)
    batch = O.synthetic_batch(T, B, A, seed=5, done_p=0.2)
    L = R.ReferenceLearner(A, state_dict=params, reward_clipping=clip)
    p0 = {k: v.clone() for k, v in params.items()}
    opt = O.new_opt_state(params)
    for step in range(2):
        st = L.learn(batch)
        ref = O.learn_step(p0, opt, batch, dict(reward_clipping=clip), use_autograd=True)
        for k in ('pg_loss', 'baseline_loss', 'entropy_loss', 'total_loss'):
            assert abs(st[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k])), (step, k)
        assert abs(st['grad_norm'] - ref['grad_norm']) <= 1e-4 * ref['grad_norm']
        sd = L.model.state_dict()
        for k in O.PARAM_ORDER:
            assert float((sd[k] - p0[k]).abs().max()) <= 2e-6, (step, k)


def test_bench_reference_arm_runs_the_reference_modules():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_mod2', os.path.join(root, 'bench.py'))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    cpu = b._CpuLearner(3, 2, 6)
    assert cpu.kind == 'reference'
    st = cpu.step()
    assert 'total_loss' in st
    assert b.workload_config(20, 32, 6, 1) == b.workload_config(20, 32, 6, 1)
