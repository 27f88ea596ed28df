"""phase stamps of conv1's weight-gradient kernel (CTA 0).  Needs a diagnostics build:
    SRL_DEFINES=SRL_WGRAD_STAMP python -m scalerl_b200.build && python tests/diag/diag_wgrad.py 20 32 ; python -m scalerl_b200.build"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
from oracle import impala_oracle as O           # noqa: E402
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams   # noqa: E402

T, B = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20, 32)
A = 6
L = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A), init_state_dict=O.init_params(A, seed=1), process_group=False)
L.use_graph = False
batch = {k: v.cuda() for k, v in O.synthetic_batch(T, B, A, seed=2).items()}
for i in range(3):
    print(f'--- step {i}', flush=True)
    L.learn(batch)
    torch.cuda.synchronize()
