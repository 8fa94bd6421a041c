"""Multi-GPU execution of one batch: MPC instances are independent (own start
state, goal, objects; shared immutable configuration), so the batch is cut
into contiguous per-rank slices, every rank solves its slice on its own GPU
with no exchange during the solve, and ONE all-gather assembles the converged
trajectories.  One process per GPU over torch.distributed
(NCCL on GPUs; gloo in the CPU tests of the host logic)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of `total` instances owned by `rank`;
    remainders go to the lowest ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(tensor, world: int, rank: int):
    lo, hi = shard_bounds(tensor.shape[0], world, rank)
    return tensor[lo:hi]


def all_gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Gather per-rank result slices (possibly uneven) into the full batch."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty((total, *local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, *local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: sizes[rank]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def sharded_solve(solve_fn, x0, target, body_params=None, group=None):
    """Solve a replicated batch description cooperatively.

    `solve_fn(x0, target, body) -> dict(X, U, status)` is the per-rank solver
    (`BatchedMPC.solve_device` on a GPU).  Every rank passes the same full
    batch tensors; returns the full X, U, status on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    total = x0.shape[0]
    out = solve_fn(shard(x0, world, rank).contiguous(), shard(target, world, rank).contiguous(),
                   None if body_params is None else shard(body_params, world, rank).contiguous())
    return {k: all_gather_batch(out[k], total, group) for k in ("X", "U", "status")}
