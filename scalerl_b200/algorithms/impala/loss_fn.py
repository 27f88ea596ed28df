"""Drop-in for scalerl/algorithms/impala/loss_fn.py (compute_baseline_loss / compute_entropy_loss /
compute_policy_gradient_loss, loss_fn.py:5-23).  On the learner hot path these three reductions are
fused with V-trace and the head gradients in one kernel (ops.impala_loss_and_head_grads); the stand-alone
functions below are thin forward-only views of that kernel for code that calls them individually."""
import torch

from ... import ops as _ops


def _tail(logits, actions=None, advantages=None):
    T, B, A = logits.shape
    dev = logits.device
    tl = torch.cat([logits, torch.zeros(1, B, A, device=dev)], 0)
    z = torch.zeros(T + 1, B, device=dev)
    act = torch.zeros(T + 1, B, dtype=torch.int64, device=dev)
    if actions is not None:
        act[1:] = actions
    return tl, z, act


@torch.no_grad()
def compute_entropy_loss(logits: torch.Tensor) -> torch.Tensor:
    """loss_fn.py:9-13: sum p log p (the negative entropy)."""
    tl, z, act = _tail(logits)
    out = _ops.impala_loss_and_head_grads(tl, tl, z, act, z, torch.ones_like(z, dtype=torch.bool), entropy_cost=1.0,
                                          baseline_cost=0.0, reward_clipping='none')
    return out['losses'][2]


@torch.no_grad()
def compute_baseline_loss(advantages: torch.Tensor) -> torch.Tensor:
    """loss_fn.py:5-6: 0.5 * sum(adv^2) -- a plain reduction; kept in torch (not on the fused path)."""
    return 0.5 * torch.sum(advantages ** 2)


@torch.no_grad()
def compute_policy_gradient_loss(logits, actions, advantages):
    """loss_fn.py:16-23: sum(-log pi(a) * adv)."""
    from .vtrace import action_log_probs
    return torch.sum(-action_log_probs(logits, actions) * advantages)
