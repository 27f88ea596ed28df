import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import impala_oracle as O
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams
T, B, A = 20, 32, 6
L = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A), init_state_dict=O.init_params(A, seed=1), process_group=False)
batch = {k: v.cuda() for k, v in O.synthetic_batch(T, B, A, seed=5).items()}
with torch.cuda.stream(torch.cuda.Stream()):
    for it in range(4):
        L.forward_backward(batch); torch.cuda.synchronize()
