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


class PipelinedSolveGather:
    """Back-to-back sharded solves with the result exchange off the critical path.

    Every rank solves its slice into one packed buffer `[X | U]` (so ONE all-gather moves a step's trajectories) and
    the all-gather of step s runs on a side stream while the solve kernel of step s+1 already occupies the SMs; two
    buffers alternate.  `finish()` joins the side stream (call it before reading results or stopping a timer)."""

    def __init__(self, mpc, batch, group=None):
        self.mpc, self.B, self.group = mpc, batch, group
        self.world = dist.get_world_size(group)
        dev, dt = torch.device("cuda", torch.cuda.current_device()), mpc.torch_dtype
        self.nX, self.nU = (mpc.N + 1) * mpc.nx, mpc.N * mpc.nu
        self.local = [torch.empty(batch * (self.nX + self.nU), dtype=dt, device=dev) for _ in range(2)]
        self.full = [torch.empty(self.world * batch * (self.nX + self.nU), dtype=dt, device=dev) for _ in range(2)]
        self.status = torch.empty(batch, dtype=torch.int32, device=dev)
        self.stats = torch.empty((batch, 8), dtype=dt, device=dev)
        self.comm = torch.cuda.Stream(device=dev)
        self.solved = [torch.cuda.Event() for _ in range(2)]
        self.gathered = [None, None]
        self.step_index = 0

    def views(self, i):
        """(X [B, N+1, nx], U [B, N, nu]) views of local buffer i."""
        buf, B = self.local[i], self.B
        return (buf[: B * self.nX].view(B, self.mpc.N + 1, self.mpc.nx), buf[B * self.nX:].view(B, self.mpc.N, self.mpc.nu))

    def gathered_views(self, i):
        """Per-rank (X, U) views of gathered buffer i: X [world, B, N+1, nx], U [world, B, N, nu]."""
        per = self.B * (self.nX + self.nU)
        full = self.full[i].view(self.world, per)
        return (full[:, : self.B * self.nX].reshape(self.world, self.B, self.mpc.N + 1, self.mpc.nx),
                full[:, self.B * self.nX:].reshape(self.world, self.B, self.mpc.N, self.mpc.nu))

    def step(self, x0, target, body=None):
        i = self.step_index & 1
        self.step_index += 1
        cur = torch.cuda.current_stream()
        if self.gathered[i] is not None:
            cur.wait_event(self.gathered[i])          # the gather that last read this buffer has finished
        X, U = self.views(i)
        self.mpc.solve_device(x0, target, body, X=X, U=U, status=self.status, stats=self.stats)
        self.solved[i].record(cur)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.solved[i])
            dist.all_gather_into_tensor(self.full[i], self.local[i], group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm)
            self.gathered[i] = ev
        return i

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.comm)


class _DeviceBuffer:
    """A raw device allocation presented to torch through `__cuda_array_interface__` (torch.as_tensor wraps it
    without a copy and keeps this object alive)."""

    def __init__(self, ptr, shape, typestr):
        self.ptr = ptr
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


class FusedSolveGather:
    """Sharded solves whose result exchange is done BY THE SOLVE KERNEL: every rank owns a gathered buffer
    [world * B] for X and U; the buffers of all ranks are mapped into every process (CUDA IPC through the library's
    own `ub_gather_alloc` / `ub_gather_open`, opened with the accessing device current so that peer access is on) and
    `ub_set_gather_targets` makes the kernel's epilogue store each solved instance into its row of every peer's
    buffer (and, through the X / U arguments, of its own) while the rest of the batch is still being solved.  No
    collective kernel, no copy engine work, nothing competing with the persistent solve grid for SMs.  Two buffer
    sets alternate so that step s + 1 never overwrites rows a consumer of step s may still be reading;
    `finish()` = local stream synchronisation + one barrier, after which `gathered_views(i)` is complete on every
    rank."""

    def __init__(self, mpc, batch, group=None):
        import ctypes as C
        from . import bindings as Bd
        self.mpc, self.B, self.group = mpc, batch, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world - 1 > Bd.UB_MAX_GATHER:
            raise ValueError(f"at most {Bd.UB_MAX_GATHER} peers")
        dev, dt = torch.device("cuda", torch.cuda.current_device()), mpc.torch_dtype
        W, B, lib = self.world, batch, mpc.lib
        esz = torch.empty((), dtype=dt).element_size()
        typestr = "<f4" if esz == 4 else "<f8"
        nX, nU = W * B * (mpc.N + 1) * mpc.nx, W * B * mpc.N * mpc.nu
        self._u_off = (nX * esz + 255) // 256 * 256
        nbytes = self._u_off + nU * esz
        self._own, self._opened, handles = [], [], []
        self.Xfull, self.Ufull = [], []
        for _ in range(2):
            ptr, h = C.c_void_p(), C.create_string_buffer(Bd.UB_IPC_HANDLE_BYTES)
            Bd.check(lib.ub_gather_alloc(nbytes, C.byref(ptr), h))
            self._own.append(ptr.value)
            handles.append(h.raw)
            self.Xfull.append(torch.as_tensor(_DeviceBuffer(ptr.value, (W * B, mpc.N + 1, mpc.nx), typestr), device=dev))
            self.Ufull.append(torch.as_tensor(_DeviceBuffer(ptr.value + self._u_off, (W * B, mpc.N, mpc.nu), typestr), device=dev))
        self.status = torch.empty(B, dtype=torch.int32, device=dev)
        self.stats = torch.empty((B, 8), dtype=dt, device=dev)
        everyone = [None] * W
        dist.all_gather_object(everyone, handles, group=group)
        self._targets = []
        for i in range(2):
            xs, us = [], []
            for r in range(W):
                if r == self.rank:
                    continue
                ptr = C.c_void_p()
                Bd.check(lib.ub_gather_open(everyone[r][i], C.byref(ptr)))
                self._opened.append(ptr.value)
                xs.append(ptr.value)
                us.append(ptr.value + self._u_off)
            self._targets.append(((C.c_void_p * len(xs))(*xs), (C.c_void_p * len(us))(*us), len(xs)))
        self.step_index = 0
        dist.barrier(group)

    def close(self):
        """Unmap the peers' buffers and, after a barrier, free the own ones (views handed out become invalid)."""
        lib = self.mpc.lib
        torch.cuda.synchronize()
        for p in self._opened:
            lib.ub_gather_close(p)
        self._opened = []
        dist.barrier(self.group)
        self.Xfull, self.Ufull = [], []
        for p in self._own:
            lib.ub_gather_free(p)
        self._own = []

    def step(self, x0, target, body=None):
        from . import bindings as Bd
        i = self.step_index & 1
        self.step_index += 1
        lo = self.rank * self.B
        xs, us, n = self._targets[i]
        Bd.check(self.mpc.lib.ub_set_gather_targets(self.mpc.handle, n, xs, us, lo))
        try:
            self.mpc.solve_device(x0, target, body, X=self.Xfull[i][lo:lo + self.B], U=self.Ufull[i][lo:lo + self.B],
                                  status=self.status, stats=self.stats)
        finally:
            Bd.check(self.mpc.lib.ub_set_gather_targets(self.mpc.handle, 0, None, None, 0))
        return i

    def gathered_views(self, i):
        """(X [world, B, N+1, nx], U [world, B, N, nu]) of buffer set i."""
        return (self.Xfull[i].view(self.world, self.B, self.mpc.N + 1, self.mpc.nx),
                self.Ufull[i].view(self.world, self.B, self.mpc.N, self.mpc.nu))

    def finish(self):
        torch.cuda.current_stream().synchronize()
        dist.barrier(self.group)
