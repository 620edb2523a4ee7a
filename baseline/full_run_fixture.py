"""TEST INFRASTRUCTURE ONLY -- records complete training runs (params.py:45 epochs = 5000, early stopping as in
functions.py:782-800) of the UNMODIFIED reference on its CUDA-eager path and of the drop-in, all driven by the
unmodified main.py (baseline/run_main.py), as the fixture tests/golden/ref_cuda_trajectory_4061.npz.

    python baseline/full_run_fixture.py --ref-runs 3 --ours-runs 2 --out tests/golden/ref_cuda_trajectory_4061.npz

The runs execute concurrently on one GPU (the reference's step is launch-bound: ~125 ms per epoch of which the GPU is
busy a fraction).  Same seed (65535), image (strawberry.jpeg) and parameter ID (4061) everywhere; the reference's
scatter-adds are atomics, so its runs differ from each other -- which is why several are recorded: their spread is the
yardstick for the drop-in's difference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-runs", type=int, default=3)
    ap.add_argument("--ours-runs", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=5000)
    ap.add_argument("--max-seconds", type=float, default=1500.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "ref_cuda_trajectory_4061.npz"))
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="gngf_full_")
    jobs = [("reference", i) for i in range(a.ref_runs)] + [("dropin", i) for i in range(a.ours_runs)]
    procs = []
    for impl, i in jobs:
        out = os.path.join(tmp, f"{impl}_{i}.npz")
        cmd = [sys.executable, os.path.join(HERE, "run_main.py"), "--impl", impl, "--epochs", str(a.epochs), "--out", out,
               "--max-seconds", str(a.max_seconds)]
        log = open(os.path.join(tmp, f"{impl}_{i}.log"), "w")
        procs.append((impl, i, out, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT,
                                                     env=dict(os.environ, WANDB_MODE="disabled"))))
    arrays, summary = {}, {"reference": [], "dropin": []}
    for impl, i, out, p in procs:
        rc = p.wait()
        if rc != 0:
            print(open(os.path.join(tmp, f"{impl}_{i}.log")).read()[-3000:])
            raise SystemExit(f"{impl} run {i} failed (rc {rc})")
        z = np.load(out)
        info = json.loads(str(z["info"]))
        key = "ref" if impl == "reference" else "ours"
        arrays[f"{key}_psnr_{i}"] = z["psnr"].astype(np.float32)
        arrays[f"{key}_loss_{i}"] = z["loss"].astype(np.float32)
        arrays[f"{key}_mse_{i}"] = z["mse"].astype(np.float32)
        summary[impl].append({"epochs": int(len(z["psnr"])), "best_psnr": float(z["psnr"].max()),
                              "final_psnr": float(z["psnr"][-1]), "min_mse": float(z["mse"].min()),
                              "stopped_by_time_budget": bool(info["stopped_by_time_budget"]),
                              "ms_per_epoch_median": float(np.median(z["sec"]) * 1e3), "wall_s": info["wall_s"]})
        print(impl, i, summary[impl][-1], flush=True)
    np.savez_compressed(a.out, summary=json.dumps(summary), ref_runs=a.ref_runs, ours_runs=a.ours_runs, **arrays)
    print("wrote", a.out)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
