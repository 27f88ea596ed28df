"""Drop-in for scalerl/algorithms/impala/loss_fn.py: ``compute_baseline_loss`` / ``compute_entropy_loss`` /
``compute_policy_gradient_loss`` (loss_fn.py:5-23) with the same arguments and values, as ``torch.autograd.Function``s over the
sm_100a row kernels -- so the reference's ``total_loss.backward()`` (impala_atari.py:330,343) keeps working when only this
import is swapped (stage 1).  On the full learner path (B200ImpalaLearner) these three reductions, V-trace
and the closed-form head gradients are ONE fused kernel (ops.impala_loss_and_head_grads); parity of both against the
reference's modules: tests/test_gpu_shims.py."""
from ...ops import compute_baseline_loss, compute_entropy_loss, compute_policy_gradient_loss  # noqa: F401
