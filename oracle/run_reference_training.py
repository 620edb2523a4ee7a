"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference `main.py` for a bounded number of epochs on CPU
and records the per-epoch trajectory (PSNR on the int image, losses, collisions) as a golden fixture.

    python oracle/run_reference_training.py --epochs 8 --out tests/golden/ref_trajectory_4061.npz

Recipe = SURVEY.md section 8c: writable cwd with an `images/` symlink, stub matplotlib, WANDB disabled, CPU
device override, `functions.epochs` overridden, then runpy of the reference's main.py with
`-f strawberry.jpeg -s 4061 -e 4061`.  `train_step` and `calc_psnr` are wrapped by pass-through observers
only (no behaviour change).
"""
import argparse
import os
import runpy
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--param-id", type=int, default=4061)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_trajectory_4061.npz"))
    args = ap.parse_args()
    out_path = os.path.abspath(args.out)
    torch.set_num_threads(args.threads)
    os.environ["WANDB_MODE"] = "disabled"

    ref = ref_shim.load_reference("cpu")
    ref_shim.set_flag(ref, "epochs", args.epochs)
    ref_shim.set_flag(ref, "should_save_params", False)   # params.py:2 -- do not write weight files
    rec = {"psnr": [], "loss": [], "mse": [], "kl": [], "coll_loss": [], "collisions": [], "sec": []}
    init = {}
    real_step, real_psnr, real_opt = ref.functions.train_step, ref.functions.calc_psnr, ref.functions.get_optimizer

    def get_opt(net, *a, **k):          # called right after the model is constructed (functions.py:568)
        for key, v in net.state_dict().items():
            if not key.startswith("_batch_norm"):
                init["init." + key] = v.detach().cpu().numpy().copy()
        return real_opt(net, *a, **k)

    def step(*a, **k):
        if "shuffled_indices" not in init:
            init["shuffled_indices"] = k["shuffled_indices"].cpu().numpy().copy()
            init["reordered_indices"] = k["reordered_indices"].cpu().numpy().copy()
        t0 = time.time()
        r = real_step(*a, **k)
        rec["sec"].append(time.time() - t0)
        rec["loss"].append(r[0]); rec["collisions"].append(r[2].cpu().numpy()); rec["mse"].append(r[5])
        rec["kl"].append(np.asarray(r[6])); rec["coll_loss"].append(np.asarray(r[7]))
        return r

    def psnr(pred, target):
        v = real_psnr(pred, target)
        rec["psnr"].append(v)
        print(f"epoch {len(rec['psnr']) - 1}: psnr {v:.6f} loss {rec['loss'][-1]:.6f} ({rec['sec'][-1]:.0f}s)", flush=True)
        return v

    ref.functions.train_step = step
    ref.functions.calc_psnr = psnr
    ref.functions.get_optimizer = get_opt
    # wandb is disabled; wandb.Image() would still try to convert the stub matplotlib figures (functions.py:752)
    ref.functions.wandb.Image = lambda *a, **k: None
    # main.py star-imports from these module objects, which are already in sys.modules
    sys.path.insert(0, ref_shim.REFERENCE_DIR)
    work = tempfile.mkdtemp(prefix="gngf_ref_run_")
    os.symlink(os.path.join(ref_shim.REFERENCE_DIR, "images"), os.path.join(work, "images"))
    os.chdir(work)
    sys.argv = ["main.py", "-f", "strawberry.jpeg", "-s", str(args.param_id), "-e", str(args.param_id)]
    runpy.run_path(os.path.join(ref_shim.REFERENCE_DIR, "main.py"), run_name="__main__")
    np.savez_compressed(out_path, param_id=args.param_id, threads=args.threads,
                        **{k: np.asarray(v) for k, v in rec.items()}, **init)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
