"""torchrun --nproc-per-node N tools/gpu_fused_gather_check.py — the result gather fused into the solve kernel
(FusedSolveGather: P2P stores from the epilogue) against the NCCL all-gather (PipelinedSolveGather): identical gathered
buffers on every rank, and the time of K back-to-back steps with either."""
import os
import sys
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload  # noqa: E402
from upright_b200.distributed import FusedSolveGather, PipelinedSolveGather  # noqa: E402
from upright_b200.engine import BatchedMPC  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name, B, K = "cfg2_thing_demo", int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 10
desc, meta = workload.load(name)
mpc = BatchedMPC(desc, "f32")
ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))  # noqa: E731
sets = []
for s in range(3):
    b = workload.sample_batch(name, desc, meta, B, 77 + 1000 * rank + s, ee)
    sets.append(tuple(torch.tensor(b[k], dtype=torch.float32, device=dev) for k in ("x0", "target", "body_params")))
nccl = PipelinedSolveGather(mpc, B)
fused = FusedSolveGather(mpc, B)
ok = True
for s in range(3):
    i = nccl.step(*sets[s]); nccl.finish(); torch.cuda.synchronize(); dist.barrier()
    Xn, Un = nccl.gathered_views(i)
    j = fused.step(*sets[s]); fused.finish()
    Xf, Uf = fused.gathered_views(j)
    same = bool(torch.equal(Xn, Xf) and torch.equal(Un, Uf))
    ok = ok and same
    if rank == 0:
        print(f"step {s}: gathered X/U identical on rank 0: {same}; rows from other ranks non-zero: "
              f"{bool((Xf[(rank + 1) % world].abs().sum() > 0).item())}", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
for label, pipe in (("nccl all-gather on a side stream", nccl), ("fused into the solve kernel", fused)):
    for s in range(3):
        pipe.step(*sets[s % 3])
    pipe.finish(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        pipe.step(*sets[s % 3])
    pipe.finish()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:36s}: {t.item() / K:8.3f} ms / step, {world * B * K / t.item() * 1e3:10.0f} solves/s over {world} GPUs", flush=True)
if rank == 0:
    print("ALL RANKS IDENTICAL" if flag.item() == 1.0 else "MISMATCH")
dist.barrier()
dist.destroy_process_group()
