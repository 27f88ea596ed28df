"""diagnostic: conv bias gradients of the device step vs the bf16-emulating oracle at several sizes"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import impala_oracle as O
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams

for (T, B) in [(5, 6), (7, 19), (9, 24), (20, 32)]:
    A = 6
    params = O.init_params(A, seed=4)
    L = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A), init_state_dict=params, process_group=False)
    batch = O.synthetic_batch(T, B, A, seed=20, done_p=0.1)
    opt = O.new_opt_state(params, 'rmsprop')
    ref = O.learn_step(params, opt, batch, dict(optimizer='rmsprop'), emulate_bf16=True)
    L.learn({k: v.cuda() for k, v in batch.items()}, use_graph=False)
    for k in ('conv1.bias', 'conv2.bias', 'conv3.bias', 'conv1.weight'):
        g, r = L.grads[k].cpu().flatten(), ref['grads'][k].flatten()
        e = float((g - r).norm() / r.norm())
        print(T, B, k, 'rel_l2 %.4f' % e, 'ratio', [round(float(x), 3) for x in (g / r)[:6]])
