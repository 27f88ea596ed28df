"""Drop-in for scalerl/algorithms/impala/vtrace.py (same function names, argument meaning, namedtuples),
backed by the sm_100a kernels.  ``from scalerl_b200.algorithms.impala.vtrace import from_logits`` replaces
``from scalerl.algorithms.impala.vtrace import from_logits`` at the call site impala_atari.py:15,310."""
import torch

from ...ops import (VTraceFromLogitsReturns, VTraceReturns, from_importance_weights, from_logits)  # noqa: F401
from ... import ops as _ops


@torch.no_grad()
def action_log_probs(policy_logits: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """vtrace.py:31-40 -- log pi(a) for [.., A] logits; computed by the from_logits kernel (on-policy trick:
    behaviour == target gives log_rho = 0 and returns the action log-prob)."""
    shp = actions.shape
    lg = policy_logits.reshape(1, -1, policy_logits.shape[-1])
    a = actions.reshape(1, -1)
    z = torch.zeros(1, a.shape[1], device=lg.device)
    r = _ops.from_logits(lg, lg, a, z, z, z, z[0])
    return r.target_action_log_probs.view(shp)
