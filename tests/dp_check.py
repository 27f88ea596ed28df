"""Multi-GPU data-parallel check (run with torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_check.py

Every rank learns on its column shard (gradient SUM over the ranks inside the step -- peer-memory apply kernel, or NCCL --
captured in the CUDA graph); rank 0 also runs a single-GPU learner on the FULL batch.  Every step (eager warm-up, capture,
replays) the sharded gradients / weights must equal the full-batch ones up to fp32 summation order -- the property
SURVEY.md §8e asks for.  After each comparison the full-batch learner is re-synchronised to the sharded state: one bf16
rounding flip of a weight or a ReLU mask turns 1e-7 into 1e-4 a step later (seen at N = 8), which says nothing about the
reduction."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import impala_oracle as O                       # noqa: E402  (checker/input generator only)
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams   # noqa: E402
from scalerl_b200 import parallel as par                    # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    T, A, B = 6, 6, 4 * world
    params = O.init_params(A, seed=11)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(T, B, A, seed=5, done_p=0.1).items()}
    mine = {k: v.contiguous() for k, v in par.shard_columns(batch, rank, world).items()}
    for optimizer in ('rmsprop', 'adam'):
        shard = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B // world, num_actions=A, optimizer=optimizer),
                                  init_state_dict=params)
        full = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A, optimizer=optimizer),
                                 init_state_dict=params, process_group=False) if rank == 0 else None
        path = (f'peer memory (fused reduce+clip+optimizer kernel; {getattr(shard, "dp_path", "?")})' if shard._peers is not None
                else 'NCCL all-reduce')
        worst = 0.0
        for step in range(5):
            s = shard.learn(mine)
            if rank == 0:
                f = full.learn(batch)
                rel = float((shard.flat_params - full.flat_params).norm() / full.flat_params.norm())
                gl = float((shard.flat_grads - full.flat_grads).norm() / full.flat_grads.norm())
                worst = max(worst, rel)
                print(f'{optimizer} step {step}: total_loss shard-sum {s["total_loss"]:.5f} full {f["total_loss"]:.5f} | grad rel-L2 {gl:.2e} | '
                      f'param rel-L2 {rel:.2e} | grad_norm {s["grad_norm"]:.4f} vs {f["grad_norm"]:.4f} | graphs {len(shard._graphs)} '
                      f'(nodes per step: {[len(g) for g in shard._graphs.values()]})', flush=True)
                assert abs(s['grad_norm'] - f['grad_norm']) <= 1e-4 * f['grad_norm'] and gl < 1e-4 and rel < 1e-5
                full.flat_params.copy_(shard.flat_params)                      # next step starts from identical state
                full.opt_state0.copy_(shard.opt_state0)
                if full.opt_state1 is not None:
                    full.opt_state1.copy_(shard.opt_state1)
        # replicas must stay bit-identical: compare an integer checksum of the weights over the ranks
        chk = shard.flat_params.view(torch.int32).to(torch.int64).sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        if rank == 0:
            assert all(int(c) == int(allc[0]) for c in allc), [int(c) for c in allc]
            assert worst < 1e-4, worst
            print(f'{optimizer}: gradient path = {path}; replicas bit-identical on {world} ranks', flush=True)
        shard.release_graphs()
    if rank == 0:
        print('DP CHECK OK', flush=True)
    dist.barrier(device_ids=[local])
    os._exit(0)


if __name__ == '__main__':
    main()
