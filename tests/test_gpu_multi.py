"""Multi-GPU data-parallel correctness, collected by pytest: spawns ``torchrun tests/dp_check.py`` on
every visible power-of-two GPU count >= 2 (skipped on a single-GPU box).  dp_check compares, every step, the column-sharded
learners (gradient SUM inside the step: peer-memory apply kernel, or NCCL) against a full-batch learner on rank 0 and checks
that the replicas stay bit-identical -- for RMSprop and Adam.  The log of the largest run is kept under gpurun_out/."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _counts():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n]


@pytest.mark.parametrize('world', [2, 4, 8])
@pytest.mark.parametrize('fused', ['nvls', 'peer_memory', 'nccl'])
def test_dp_check(world, fused):
    if world not in _counts():
        pytest.skip(f'{world} GPUs not visible')
    env = dict(os.environ)
    env['SRL_DP_FUSED'] = '0' if fused == 'nccl' else '1'
    env['SRL_DP_NVLS'] = '1' if fused == 'nvls' else '0'
    env['MASTER_ADDR'] = '127.0.0.1'
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dp_check.py')]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    out = r.stdout + '\n' + r.stderr
    d = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f'dp_check_n{world}_{fused}.log'), 'w') as f:
        f.write(out)
    assert r.returncode == 0 and 'DP CHECK OK' in out, out[-4000:]
    want = {'nvls': 'nvls multimem', 'peer_memory': 'peer loads', 'nccl': 'NCCL all-reduce'}[fused]
    if fused == 'nvls' and 'nvls multimem' not in out:
        pytest.skip('no NVLS multicast mapping on this box (torch symmetric memory reported multicast_ptr = 0): ran on peer loads')
    assert want in out, out[-2000:]


def test_multi_learner_trainer_shares_one_ring(tmp_path):
    """ImpalaTrainer(num_learners=2): two learner PROCESSES (one per GPU) dequeue from ONE pinned trajectory ring fed by the same
    actors, gradients SUM-reduced inside the step; replicas end bit-identical and rank 0 published the weights.
    Runs in a child interpreter: the trainer forks before CUDA is touched, which a pytest process that already used CUDA cannot do."""
    if 2 not in _counts():
        pytest.skip('2 GPUs not visible')
    code = f"""
import sys, json, torch
sys.path.insert(0, {ROOT!r})
from scalerl_b200.algorithms.impala.impala_atari import ImpalaArguments, ImpalaTrainer
a = ImpalaArguments(num_actors=3, num_learners=2, batch_size=4, rollout_length=5, num_buffers=24, total_steps=2 * 4 * 5 * 6, output_dir={str(tmp_path)!r})
t = ImpalaTrainer(a)
w0 = t.actor_model.state_dict()['fc.weight'].clone()
out = t.train()
ck = torch.load({str(tmp_path)!r} + '/impala/model.tar', weights_only=False)
same = all(torch.equal(ck['model_state_dict'][k], v) for k, v in t.actor_model.state_dict().items())
print('RESULT ' + json.dumps(dict(steps=out['steps'], learners=out['learners'], checks=out['replica_checksums'], version=out['weights_version'],
      loss=out['total_loss'], moved=not torch.equal(w0, t.actor_model.state_dict()['fc.weight']), actor_equals_checkpoint=same, path=out['grad_path'])))
"""
    env = dict(os.environ, SRL_LEARNER_TIMEOUT_S='150')
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    out = r.stdout + '\n' + r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
    assert r.returncode == 0 and line, out[-4000:]
    import json
    res = json.loads(line[0][7:])
    assert res['steps'] == 2 * 4 * 5 * 6 and res['learners'] == 2
    assert len(set(res['checks'])) == 1, res                       # replicas bit-identical
    assert res['version'] == 6 and res['moved'] and res['actor_equals_checkpoint'], res
