"""Per-call device times of one default-workload step with the persistent kernels limited to GNGF_DEBUG_SM_LIMIT SMs.

    for n in 148 111 74; do GNGF_DEBUG_SM_LIMIT=$n python profiles/sm_limit_probe.py; done

A kernel bound inside the SM (tensor pipe, epilogue) slows down in proportion; one bound by what the SMs share (L2 ->
SM bandwidth) does not.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from collision_handling_in_instantngp_b200 import _lib, ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
torch.cuda.set_device(0)
R = bench.Runner(torch, None, name, dict(bench.WORKLOADS[name]), 0, 1, torch.device("cuda", 0))
R.step(R.x_dev, R.y_dev)
prof = bench.CallProfiler(torch)
_lib.PROFILER = prof
ops.CONCURRENT = False
for _ in range(2):
    R.step(R.x_dev, R.y_dev)
_lib.PROFILER = None
agg = {}
for (k, _), (t, n) in prof.summary().items():
    a = agg.setdefault(k, [0.0, 0])
    a[0] += t
    a[1] += n
print(json.dumps({"sm_limit": os.environ.get("GNGF_DEBUG_SM_LIMIT"),
                  "ms_per_step": {k: round(v[0] / 2, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:6]}}))
