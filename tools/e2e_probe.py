import sys, time, numpy as np
sys.path.insert(0, '.')
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
desc, meta = workload.load("cfg2_thing_demo")
mpc = BatchedMPC(desc, "f32")
ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))
b = workload.sample_batch("cfg2_thing_demo", desc, meta, 4096, 1, ee)
res = None
for i in range(6):
    t0 = time.perf_counter()
    res = mpc.solve(b["x0"], b["target"], b["body_params"], out=res)
    print("python total %.3f ms" % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
