"""Workload for the Nsight Compute captures (tools/ncu_capture.sh): two warm-up learner steps, then -- inside the NVTX range
"capture" -- one eager learner step at T=20,B=32 (every kernel of the default path) and the stand-alone V-trace kernels at the
sizes BASELINE.json quotes (T=20,B=512 scan / sequential; T=100,B=128 scan)."""
import os
import sys

os.environ.setdefault('SRL_NO_GRAPH', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
from oracle import impala_oracle as O           # noqa: E402  (input generator only)
from scalerl_b200 import ops                    # noqa: E402
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams   # noqa: E402


def main():
    T, B, A = 20, 32, 6
    L = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A), process_group=False, use_graph=False, seed=0)
    batches = [{k: v.cuda() for k, v in O.synthetic_batch(T, B, A, seed=s).items()} for s in range(3)]
    g = torch.Generator(device='cuda').manual_seed(0)
    r = lambda *s: torch.randn(*s, device='cuda', generator=g)
    vt = {(t, b): [r(t, b) * .3, torch.full((t, b), 0.99, device='cuda'), r(t, b), r(t, b), r(b)] for (t, b) in ((20, 512), (100, 128))}
    for i in range(2):
        L.learn(batches[i], sync_stats=False)
        for (t, b), a in vt.items():
            ops.from_importance_weights(*a, variant=1)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push('capture')
    L.learn(batches[2], sync_stats=False)
    ops.from_importance_weights(*vt[(20, 512)], variant=1)
    ops.from_importance_weights(*vt[(20, 512)], variant=0)
    ops.from_importance_weights(*vt[(100, 128)], variant=1)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print('ncu_step done', float(L._losses[3]))


if __name__ == '__main__':
    main()
