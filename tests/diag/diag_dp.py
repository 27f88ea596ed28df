"""torchrun diagnostic: phase times of the fused peer-memory apply kernel (SRL_DP_DEBUG=1) at the bench workload"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from oracle import impala_oracle as O
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
T, B, A = 20, 32, 6
L = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A), init_state_dict=O.init_params(A, seed=1))
batch = {k: v.cuda() for k, v in O.synthetic_batch(T, B, A, seed=5 + rank).items()}
with torch.cuda.stream(torch.cuda.Stream()):
    for it in range(6):
        L.learn(batch, sync_stats=False)
    torch.cuda.synchronize()
L.release_graphs()
dist.barrier(device_ids=[local])
os._exit(0)
