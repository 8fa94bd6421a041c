#!/bin/bash
# One GPU-box call: parity tests, the headline bench, the per-phase cycle counters.
# usage: gpurun -- 'bash tools/gpu_quick.sh [tag]'
tag=${1:-quick}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; tail -4 gpurun_out/pytest_${tag}.log
timeout 300 python bench.py --cpu-seconds 2 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; python - <<P
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}.json"))
    print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"iters",d["mean_qp_iterations"],"conv",d["converged_fraction"],"clk",d["clocks"])
except Exception as e:
    print("bench failed",e); print(open("gpurun_out/bench_${tag}.err").read()[-2000:])
P
timeout 120 python tools/gpu_phase_profile.py > gpurun_out/phase_${tag}.log 2>&1; cat gpurun_out/phase_${tag}.log
