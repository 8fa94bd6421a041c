#!/usr/bin/env python3
"""Benchmark of the batched waiter's-problem MPC solve (BASELINE.json metric:
MPC solves/sec, 9-DoF Thing, 1 object, 20-knot horizon).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config cfg2_thing_demo] [--batch B]

A "step" is one batched solve of B independent MPC instances (cold start, one
SQP iteration with the QP solved to its interior-point tolerance = one
`advanceMpc()` per instance).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "MPC solves/sec (9-DoF Thing, 1 object, 20-knot horizon)"
UNIT = "solves/s"


def measured_traffic(config, batch):
    """DRAM bytes per launch of the solve kernel from the committed `ncu --set full` capture of this round
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    try:
        with open(ROOT / "profiles" / "traffic.json") as f:
            t = json.load(f).get(config)
        if t and int(t.get("batch", -1)) == int(batch):
            return float(t["dram_bytes_per_launch"]), t.get("source")
    except Exception:
        pass
    return None, None


def measured_peaks():
    try:
        with open(ROOT / "MEASURED_PEAKS.json") as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def mark(self):
        return time.perf_counter()

    def wait_first(self, timeout=3.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken in [t_begin, t_end] (the timed region); nvidia-smi needs ~0.1 s to deliver
        its first sample, so the sampler is started before the warm-up."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [s for (ts, s) in self.samples if (t_begin is None or ts >= t_begin) and (t_end is None or ts <= t_end + 0.03)]
        window = "timed region"
        if not inside:
            inside, window = [s for (_, s) in self.samples], "warm-up + timed region (timed region shorter than one sample)"
        for s in inside:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def oracle_rate(desc, batch, threads, min_seconds=10.0, max_instances=None, chunk=None):
    """Time the CPU oracle (restated OCS2-equivalent, fp64) on a bounded sample."""
    import oracle
    n = min(len(batch["x0"]), max_instances or len(batch["x0"]))
    done, t0 = 0, time.perf_counter()
    chunk = chunk or max(threads * 4, 32)
    while True:
        idx = np.arange(done, done + chunk) % n
        bp = None if batch["body_params"] is None else batch["body_params"][idx]
        oracle.solve_batch(desc, batch["x0"][idx], batch["target"][idx], bp, nthreads=threads)
        done += chunk
        el = time.perf_counter() - t0
        if el >= min_seconds:
            return done / el, done, el


PER_GPU_BATCH = {   # BASELINE.json batch of a configuration on ONE GPU (cfg4: 16384 / 8, cfg5: 8192 / 8)
    "cfg1_ur10_demo": 4096, "cfg2_thing_demo": 4096, "cfg3_thing_box_arch": 4096, "cfg4_thing_obstacles2": 2048,
    "cfg5_thing_robust8": 1024,
}


FULL_BATCH_ONE_GPU = {"cfg4_thing_obstacles2": 16384, "cfg5_thing_robust8": 8192}   # BASELINE's whole batch on ONE GPU


def config_block(name, prec, world, dev, peak_tflops, seed=4321, batch=None, gather="fused"):
    """One BASELINE configuration at its per-GPU batch: kernel time (CUDA events, mean of 3 launches after a warm-up),
    iteration statistics, status counts and the arithmetic roofline fraction.  With world > 1 every rank solves its own
    shard (cfg4: 8 x 2048 = 16384, cfg5: 8 x 1024 = 8192 — BASELINE's split) and the packed results are all-gathered;
    the time is the max over ranks, the rate the global batch over it."""
    import torch
    import torch.distributed as dist
    from upright_b200 import workload
    from upright_b200.engine import BatchedMPC
    desc, meta = workload.load(name)
    mpc = BatchedMPC(desc, prec)
    dt = mpc.torch_dtype
    B = batch or PER_GPU_BATCH[name]
    rank = dist.get_rank() if world > 1 else 0
    ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))  # noqa: E731
    mg = (lambda x: mpc.eval("obstacle_avoidance", x, np.zeros((x.shape[0], mpc.nu)))) if desc.obstacles_enabled else None
    b = workload.sample_batch(name, desc, meta, B, seed + 1000 * rank, ee, margin_fn=mg)
    t = lambda a: None if a is None else torch.tensor(a, dtype=dt, device=dev)  # noqa: E731
    x0, tg, bp = t(b["x0"]), t(b["target"]), t(b["body_params"])
    nX, nU = (mpc.N + 1) * mpc.nx, mpc.N * mpc.nu
    packed = torch.empty(B * (nX + nU), dtype=dt, device=dev)
    X, U = packed[: B * nX].view(B, mpc.N + 1, mpc.nx), packed[B * nX:].view(B, mpc.N, mpc.nu)
    fused = None
    if world > 1 and gather == "fused":   # results stored into every peer's gathered buffer by the solve kernel itself
        from upright_b200.distributed import FusedSolveGather
        fused = FusedSolveGather(mpc, B)
    full = torch.empty(world * B * (nX + nU), dtype=dt, device=dev) if world > 1 and fused is None else None
    status = torch.empty(B, dtype=torch.int32, device=dev) if fused is None else fused.status
    stats = torch.empty((B, 8), dtype=dt, device=dev) if fused is None else fused.stats
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, kms = [], []
    for it in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        if fused is not None:
            fused.step(x0, tg, bp)
            fused.finish()                # stream synchronisation + barrier: the gathered buffers are complete on every rank
        else:
            mpc.solve_device(x0, tg, bp, X=X, U=U, status=status, stats=stats)
            if world > 1:
                dist.all_gather_into_tensor(full, packed)
        ev1.record()
        torch.cuda.synchronize()
        if it > 0:
            ms.append(ev0.elapsed_time(ev1))
            kms.append(mpc.last_solve_ms())
    tm = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device=dev)
    counts = torch.bincount(status.long().clamp(0, 3), minlength=4).double()
    it_sum = torch.stack((stats[:, 0].double().sum(), stats[:, 0].double().max()))
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        it_max = it_sum[1:].clone()
        dist.all_reduce(it_sum[:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(it_max, op=dist.ReduceOp.MAX)
        it_sum[1] = it_max[0]
    mean_it = float(it_sum[0].item()) / (B * world)
    step_ms = float(tm.item())
    flops = workload.algorithmic_flops_per_solve(desc, mean_it, desc.sqp_iteration) * B * world
    tf = flops / (step_ms * 1e-3) / 1e12
    traffic, src = measured_traffic(name, B)
    out = {"per_gpu_batch": B, "global_batch": B * world, "solves_per_s": B * world / (step_ms * 1e-3), "ms_per_batch": step_ms,
           "kernel_ms": float(np.mean(kms)), "mean_ipm_iterations": mean_it, "max_ipm_iterations": float(it_sum[1].item()),
           "status_counts": {"converged": int(counts[0].item()), "qp_maxiter": int(counts[1].item()),
                             "linesearch_failed": int(counts[2].item()), "nan": int(counts[3].item())},
           "algorithmic_tflops": tf, "fp32_frac": tf / (peak_tflops * world) if peak_tflops else None,
           "dram_bytes_per_launch": traffic, "dram_bytes_source": src,
           "nx": mpc.nx, "nu": mpc.nu, "dtype": prec}
    if fused is not None:
        out["gather"] = "fused into the solve kernel"
        fused.close()
    elif world > 1:
        out["gather"] = "ncclAllGather after the solve (not overlapped)"
    del mpc
    if world == 1 and batch is None and name in FULL_BATCH_ONE_GPU:
        # the 8-GPU configurations also as ONE batch on one GPU (several waves of the persistent grid instead of one)
        full_one = config_block(name, prec, world, dev, peak_tflops, seed, batch=FULL_BATCH_ONE_GPU[name])
        out["whole_baseline_batch_on_one_gpu"] = {k: full_one[k] for k in ("global_batch", "solves_per_s", "ms_per_batch", "kernel_ms",
                                                                            "mean_ipm_iterations", "max_ipm_iterations",
                                                                            "status_counts", "fp32_frac")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2_thing_demo")
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU per step")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-batch", type=int, default=None, help="reference arm: instances per step (default: the batch)")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration block")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="multi-GPU result exchange")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    # stdout carries the ONE JSON line: everything libraries print there meanwhile (NCCL's version banner ...) is
    # sent to stderr by pointing file descriptor 1 at it until the line is emitted
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    from upright_b200 import workload
    desc, meta = workload.load(args.config)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch or PER_GPU_BATCH.get(args.config, 4096)
    cores = os.cpu_count() or 1
    config = {"workload": f"{args.config}: {meta['source']}, batch={B}/GPU cold start, sqp_iteration={desc.sqp_iteration}, "
                          f"N={desc.N} knots, nx={desc.nx}, nu={desc.nu}",
              "global_batch": B * world, "per_gpu_batch": B, "seed": 1234,
              "parallelism": f"dp{world} (independent instances sharded; one NCCL all-gather of the packed [X|U] results per step, on a side stream under the next step's solve)" if world > 1 else "dp1",
              "l2": "no explicit flush: the per-step working set (workspace slots of every resident warp, rewritten every interior-point iteration) exceeds the 126 MB L2; inputs differ every step"}

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        import oracle
        oracle.build()

        def ee_fn(x):
            return np.array([oracle.fk(desc, xi)["r"] for xi in x])

        sample_n = args.ref_batch or B
        # the same per-step batch as our arm (one distinct seeded set per step would cost minutes of set-up on the
        # host for nothing: the oracle's time does not depend on it), rotated by step
        base = workload.sample_batch(args.config, desc, meta, sample_n, 1234, ee_fn)

        def one(s):
            idx = (np.arange(sample_n) + 17 * s) % sample_n
            bp = None if base["body_params"] is None else base["body_params"][idx]
            oracle.solve_batch(desc, base["x0"][idx], base["target"][idx], bp, nthreads=cores)

        for s in range(args.warmup):
            one(s)
        t0 = time.perf_counter()
        for s in range(args.warmup, args.warmup + args.steps):
            one(s)
        el = time.perf_counter() - t0
        value = sample_n * args.steps / el
        config["reference_per_step_batch"] = sample_n
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{sample_n} instances of the workload per step x {args.steps} steps, "
                                           f"oracle-CPU (restated OCS2-equivalent, fp64), {cores} host threads"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import ctypes
    import torch
    import torch.distributed as dist
    from upright_b200.bindings import check, load_library
    from upright_b200.engine import BatchedMPC

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load_library()
    mpc = BatchedMPC(desc, args.precision)
    dt = mpc.torch_dtype
    nsets = args.warmup + args.steps

    def ee_fn(x):
        return mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))

    sets = [workload.sample_batch(args.config, desc, meta, B, 1234 + 1000 * rank + s, ee_fn) for s in range(nsets)]
    dsets = []
    for b in sets:
        dsets.append(dict(x0=torch.tensor(b["x0"], dtype=dt, device=dev), target=torch.tensor(b["target"], dtype=dt, device=dev),
                          body=None if b["body_params"] is None else torch.tensor(b["body_params"], dtype=dt, device=dev)))
    X = torch.empty((B, mpc.N + 1, mpc.nx), dtype=dt, device=dev)
    U = torch.empty((B, mpc.N, mpc.nu), dtype=dt, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    stats = torch.empty((B, 8), dtype=dt, device=dev)
    pipe, gather_kind = None, None
    if world > 1:
        from upright_b200.distributed import FusedSolveGather, PipelinedSolveGather
        if args.gather == "fused":
            try:   # the solve kernel's epilogue stores every instance into the peers' gathered buffers (NVLink P2P)
                pipe, gather_kind = FusedSolveGather(mpc, B), "fused into the solve kernel (P2P stores from its epilogue, no collective kernel)"
            except Exception as e:  # noqa: BLE001 — e.g. no peer access: fall back to the NCCL collective, and say so
                print(f"[bench] fused gather unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
        if pipe is None:
            pipe = PipelinedSolveGather(mpc, B)   # packed [X | U] results, ONE all-gather per step on a side stream
            gather_kind = "ncclAllGather of the packed [X|U] results on a side stream under the next step's solve"

    def step(s):
        d = dsets[s]
        if pipe is not None:
            pipe.step(d["x0"], d["target"], d["body"])
        else:
            mpc.solve_device(d["x0"], d["target"], d["body"], X=X, U=U, status=status, stats=stats)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for s in range(args.warmup):
        step(s)
    if pipe is not None:
        pipe.finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.ub_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters, ok = [], []
    torch.cuda.synchronize()
    t_begin = sampler.mark()
    ev0.record()
    for s in range(args.warmup, nsets):
        step(s)
    if pipe is not None:
        pipe.finish()                     # every all-gather of the timed steps completes inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.ub_launch_count() - launches0
    t_end = sampler.mark()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    # iteration statistics of a few steps (outside the timed region)
    for s in range(args.warmup, min(nsets, args.warmup + 3)):
        step(s)
        torch.cuda.synchronize()
        st, ss = (pipe.stats, pipe.status) if pipe is not None else (stats, status)
        iters.append(float(st[:, 0].double().mean().item()))
        ok.append(float((ss == 0).double().mean().item()))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = B * world * args.steps / (ms_total * 1e-3)

    # end-to-end through the host-buffer C-ABI call (H2D + D2H inside)
    e2e_steps = max(3, min(args.steps, 10))
    res = None
    for s in range(2):
        res = mpc.solve(sets[s]["x0"], sets[s]["target"], sets[s]["body_params"], out=res)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        b = sets[(args.warmup + s) % nsets]
        res = mpc.solve(b["x0"], b["target"], b["body_params"], out=res)  # caller-owned output buffers, reused
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(t.item())
    esz = 8 if args.precision == "f64" else 4
    h2d = B * (mpc.nx + 3 * (mpc.N + 1) + (mpc.nb * 10 if sets[0]["body_params"] is not None else 0)) * esz
    d2h = B * ((mpc.N + 1) * mpc.nx + mpc.N * mpc.nu + 8) * esz + B * 4

    # receding-horizon closed loop on the device: 0.1 s of simulated time at the reference's
    # rates (1 kHz plant, 100 Hz replanning, warm starts, Riccati feedback policy); second, informational number
    closed_loop = None
    if world == 1:
        b0 = sets[0]
        goal = b0["target"][:, :1, :]
        kw = dict(n_steps=100, sim_dt=0.001, replan_period=0.01, body_params=b0["body_params"], log=False)
        mpc.closed_loop(b0["x0"], [0.0], goal, **dict(kw, n_steps=20))
        t0 = time.perf_counter()
        out = mpc.closed_loop(b0["x0"], [0.0], goal, **kw)
        el = time.perf_counter() - t0
        closed_loop = {"warm_solves_per_s": B * out["n_replans"] / el, "sim_steps_per_s": B * kw["n_steps"] / el,
                       "replans": out["n_replans"], "sim_steps": kw["n_steps"], "wall_s": el,
                       "status_counts": out["status_counts"].sum(axis=0).tolist(),
                       "api": "BatchedMPC.closed_loop -> ub_closed_loop (host x0 in, final state out, everything else on the device)"}

    # measured FP32 multiply-add peak of this GPU (register-resident FMA loop, ub_measure_fma_peak)
    pk = ctypes.c_double()
    check(lib.ub_measure_fma_peak(ctypes.byref(pk)))
    fp32_peak = float(pk.value)

    # every BASELINE configuration at its per-GPU batch (cfg4 / cfg5: BASELINE's 8-GPU split when world = 8)
    configs = None
    if not args.no_configs:
        configs = {}
        for name in PER_GPU_BATCH:
            try:
                configs[name.split("_")[0]] = config_block(name, args.precision, world, dev, fp32_peak, gather=args.gather)
            except Exception as e:  # noqa: BLE001 — a failing extra must not cost the headline line
                configs[name.split("_")[0]] = {"error": str(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    if gather_kind:
        config["parallelism"] = f"dp{world} (independent instances sharded, no exchange during the solve; result gather: {gather_kind})"
    peaks, peak_kind = measured_peaks()
    kms = ms_total / args.steps          # the solve kernel (+ a 4-byte memset of its work queue) is the whole step
    abytes = workload.algorithmic_bytes_per_solve(desc) * B
    achieved_gbs = abytes / (kms * 1e-3) / 1e9
    mean_iters = float(np.mean(iters))
    flops = workload.algorithmic_flops_per_solve(desc, mean_iters, desc.sqp_iteration) * B
    tflops = flops / (kms * 1e-3) / 1e12
    traffic, traffic_src = measured_traffic(args.config, B)
    roofline = {"bound": "fp32-issue", "achieved": tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tflops / fp32_peak,
                "traffic": traffic, "traffic_ratio": (traffic / abytes) if traffic else None, "traffic_source": traffic_src,
                "peak_source": "measured in this run: ub_measure_fma_peak (16 independent FMA chains per thread on every SM)",
                "kernel": "ub::solve_batch_kernel", "kernel_ms": kms, "algorithmic_flops_per_launch": flops,
                "mean_ipm_iterations": mean_iters,
                "note": "many small dependent factorisations: the binding resources are FP32/FP64 issue and shared-memory "
                        "latency (SURVEY.md §8d); the algorithmic flop count is SURVEY.md §8(d)'s dense-Riccati convention, "
                        "the kernels execute fewer (forces eliminated, A and B structured)",
                "hbm": {"achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
                        "algorithmic_bytes_per_launch": abytes, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                        "note": "compulsory I/O is ~3.7 KB per solve: HBM is not the roof by construction"}}
    cpu_seconds = args.cpu_seconds if world == 1 else min(args.cpu_seconds, 5.0)
    cpu_rate, cpu_n, cpu_el = oracle_rate(desc, sets[0], cores, min_seconds=cpu_seconds)
    cpu1_rate, cpu1_n, cpu1_el = oracle_rate(desc, sets[0], 1, min_seconds=min(3.0, cpu_seconds), chunk=8)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic", "config": config, "clocks": clocks,
            "arithmetic": "f32: Riccati recursion (matrices, factors, directions) in fp32; iterate, slack / multiplier records, "
                          "residuals, linearisation and the force block in fp64" if args.precision == "f32" else "fp64 throughout",
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "upright_b200.engine.BatchedMPC.solve -> ub_solve_batch (host double buffers in/out; inputs H2D from pinned staging, every solved instance written home by the kernel through mapped pinned memory and converted float->double by host threads behind its completion flag while the kernel runs)",
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "roofline": roofline,
            "converged_fraction": float(np.mean(ok)), "mean_qp_iterations": mean_iters}
    if configs is not None:
        line["configs"] = configs
    if closed_loop is not None:
        line["closed_loop"] = closed_loop
    line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{cpu_n} instances of the same workload in {cpu_el:.1f} s, oracle-CPU "
                                      f"(restated OCS2-equivalent, fp64, dense Riccati), {cores} host threads",
                            "single_core": {"value": cpu1_rate, "sample": f"{cpu1_n} instances in {cpu1_el:.1f} s, 1 thread"}}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
