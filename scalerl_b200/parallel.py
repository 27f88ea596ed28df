"""Host-side data-parallel logic of the learner.  No CUDA here: usable with any
torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).

The B columns of a [T+1, B] trajectory batch are independent through encoder, V-trace and loss; ranks
take contiguous column shards; gradients and the three loss sums are SUM-reduced (the reference losses are
sums over T*B, loss_fn.py:6,13,23), after which every rank applies the identical clipped update
(impala_atari.py:344-346 at global batch)."""
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous split of B columns; the first B % world ranks get one extra column"""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} not in [0, {world})')
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_columns(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """rank's column shard of every [T+1, B, ...] tensor of the batch (views; call .contiguous() before the kernels)"""
    B = next(iter(batch.values())).shape[1]
    lo, hi = shard_bounds(B, rank, world)
    return {k: v[:, lo:hi] for k, v in batch.items()}


def allreduce_sum_(flat_grads: torch.Tensor, losses: torch.Tensor = None, group=None) -> None:
    """in-place SUM all-reduce of the flat fp32 gradient buffer (1.69 M floats = 6.75 MB, one bucket) and, if
    given, of the loss scalars"""
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    if losses is not None:
        dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=group)


def broadcast_params_(flat_params: torch.Tensor, src: int = 0, group=None) -> None:
    """make every rank start from rank `src`'s weights"""
    dist.broadcast(flat_params, src=src, group=group)
