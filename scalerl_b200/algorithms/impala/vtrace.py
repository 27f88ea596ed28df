"""Drop-in for scalerl/algorithms/impala/vtrace.py (same function names, argument meaning, namedtuples),
backed by the sm_100a kernels.  ``from scalerl_b200.algorithms.impala.vtrace import from_logits`` replaces
``from scalerl.algorithms.impala.vtrace import from_logits`` at the call site impala_atari.py:15,310.
Autograd behaviour follows the reference: ``from_importance_weights`` and the ``vs`` / ``pg_advantages`` of ``from_logits``
carry no graph (vtrace.py:78), ``action_log_probs`` / ``log_rhos`` / ``target_action_log_probs`` are differentiable."""
from ...ops import (VTraceFromLogitsReturns, VTraceReturns, action_log_probs, from_importance_weights, from_logits)  # noqa: F401
