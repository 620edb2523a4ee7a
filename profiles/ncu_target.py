"""Target for the ncu captures of the default bench workload (BASELINE.json configs[3], T = 2^14, one GPU): one warm
step, then ONE profiled step of the eager module path -- the same launches bench.py times.

    ncu --set full --clock-control none --import-source on -k regex:hpd_stream_ --launch-skip 3 -c 3 \\
        -o gpurun_out/prof_r02_cfg4_t14_stream python profiles/ncu_target.py [workload]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.cuda.set_device(0)
R = bench.Runner(torch, None, name, dict(bench.WORKLOADS[name]), 0, 1, torch.device("cuda", 0))
for _ in range(steps):
    R.step(R.x_dev, R.y_dev)
torch.cuda.synchronize()
print("done", name)
