"""TEST / BENCH INFRASTRUCTURE ONLY -- runs the reference's UNMODIFIED `main.py` (from baseline/_ref, see
fetch_ref.py) and records the per-epoch trajectory.

    python baseline/run_main.py --impl dropin    --epochs 500 --out /tmp/ours.npz
    python baseline/run_main.py --impl reference --epochs 500 --out /tmp/ref.npz

`--impl reference`: `models` resolves to the reference's own models.py (its CUDA-eager ATen path: the fp32 oracle
                    SURVEY.md section 8c names for the GPU box).
`--impl dropin`   : the repository root is put AHEAD of the reference directory on sys.path, so that main.py's
                    `from models import *` (main.py:3) finds this repository's drop-in module; functions.py, utils.py,
                    params.py and main.py itself are the reference's, byte for byte.

What is NOT the reference's (none of it touches the training path; SURVEY.md section 8c lists why each is needed):
  * matplotlib is absent from the image -> stub modules (functions.py:8-10 import it; plots are diagnostics);
  * WANDB_MODE=disabled, and wandb.Image() of a stub figure returns None (functions.py:749-755);
  * `functions.epochs` (params.py:45, read at functions.py:648,670) is overridden to bound the run;
  * pass-through observers around train_step / calc_psnr / get_optimizer / net.calc_hash_collisions record values.
The working directory is a scratch directory with an `images/` link (main.py:42-47 opens ./images/<file>, and
functions.py:764-780 writes ./weights/<id>_<time>/).
"""
import argparse
import importlib
import importlib.machinery
import json
import os
import runpy
import sys
import tempfile
import time
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class StopRun(Exception):
    """Raised by the epoch observer when the wall-clock budget is spent."""


def _stub_module(name):
    m = mock.MagicMock(name=name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__name__ = name
    return m


def install_matplotlib_stub():
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
        return False
    except Exception:
        pass
    mpl = _stub_module("matplotlib")
    ticker = _stub_module("matplotlib.ticker")
    pyplot = _stub_module("matplotlib.pyplot")
    figure = _stub_module("matplotlib.figure")

    class Figure:  # used in a return annotation (functions.py:363)
        pass

    figure.Figure = Figure
    mpl.figure, mpl.ticker, mpl.pyplot = figure, ticker, pyplot
    pyplot.subplots = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    sys.modules.update({"matplotlib": mpl, "matplotlib.ticker": ticker, "matplotlib.pyplot": pyplot,
                        "matplotlib.figure": figure})
    return True


def run(impl, epochs, out, ref_dir, device="cuda", param_id=4061, filename="strawberry.jpeg", save_params=False,
        workdir=None, max_seconds=0.0, threads=0, quiet=True):
    import torch

    if not os.path.isfile(os.path.join(ref_dir, "main.py")):
        raise SystemExit(f"no reference at {ref_dir}: run baseline/fetch_ref.py where /root/reference exists")
    os.environ["WANDB_MODE"] = "disabled"
    os.environ.setdefault("WANDB_SILENT", "true")
    if threads:
        torch.set_num_threads(threads)
    stubbed = install_matplotlib_stub()

    for name in ("functions", "models", "utils", "params"):
        sys.modules.pop(name, None)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, os.path.abspath(ref_dir))]
    sys.path.insert(0, ref_dir)
    if impl == "dropin":
        sys.path.insert(0, ROOT)            # <repo>/models.py shadows the reference's models.py -- nothing else does

    real_set_default_device = torch.set_default_device
    try:
        if device == "cpu":                 # CPU sanity runs of the reference only (functions.py:49-52 forces CUDA)
            torch.set_default_device = lambda *_a, **_k: None
        functions = importlib.import_module("functions")
        utils = importlib.import_module("utils")
        models = importlib.import_module("models")
    finally:
        torch.set_default_device = real_set_default_device
    if device == "cpu":
        for m in (functions, models, utils):
            m.device = torch.device("cpu")
    models_file = os.path.realpath(models.__file__)
    if impl == "dropin":
        assert models_file.startswith(os.path.realpath(ROOT)) and not models_file.startswith(os.path.realpath(ref_dir)), models_file
    else:
        assert models_file.startswith(os.path.realpath(ref_dir)), models_file
    assert os.path.realpath(functions.__file__).startswith(os.path.realpath(ref_dir)), functions.__file__

    functions.epochs = int(epochs)
    functions.should_save_params = bool(save_params)

    rec = {"psnr": [], "loss": [], "mse": [], "kl": [], "coll_loss": [], "collisions": [], "sec": []}
    info = {"impl": impl, "device": device, "models_file": models_file, "matplotlib_stubbed": stubbed,
            "collisions_input": None, "param_id": param_id, "epochs_requested": int(epochs)}
    init = {}
    real_step, real_psnr, real_opt = functions.train_step, functions.calc_psnr, functions.get_optimizer
    t_start = time.time()

    def get_opt(net, *a, **k):              # called right after the model is constructed (functions.py:568)
        for key, v in net.state_dict().items():
            init["init." + key] = v.detach().cpu().numpy().copy()
        info["net_class"] = type(net).__module__ + "." + type(net).__name__
        real_coll = net.calc_hash_collisions

        def coll(indices):                  # what train_step hands to calc_hash_collisions (functions.py:327)
            if info["collisions_input"] is None:
                info["collisions_input"] = {"dtype": str(indices.dtype), "shape": list(indices.shape),
                                            "device": str(indices.device)}
            return real_coll(indices)

        net.calc_hash_collisions = coll
        opt = real_opt(net, *a, **k)
        info["optimizer"] = type(opt).__module__ + "." + type(opt).__name__
        return opt

    def step(*a, **k):
        if "shuffled_indices" not in init:
            init["shuffled_indices"] = k["shuffled_indices"].cpu().numpy().copy()
            init["reordered_indices"] = k["reordered_indices"].cpu().numpy().copy()
        if device == "cuda":
            torch.cuda.synchronize()
        t0 = time.time()
        r = real_step(*a, **k)
        if device == "cuda":
            torch.cuda.synchronize()
        rec["sec"].append(time.time() - t0)
        rec["loss"].append(r[0])
        rec["collisions"].append(r[2].detach().cpu().numpy())
        rec["mse"].append(r[5])
        rec["kl"].append(np.asarray(r[6]))
        rec["coll_loss"].append(np.asarray(r[7]))
        return r

    def psnr(pred, target):
        v = real_psnr(pred, target)
        rec["psnr"].append(v)
        e = len(rec["psnr"]) - 1
        if not quiet or e % 100 == 0:
            print(f"[{impl}] epoch {e}: psnr {v:.4f} loss {rec['loss'][-1]:.6f} ({rec['sec'][-1] * 1e3:.1f} ms/epoch)",
                  file=sys.stderr, flush=True)
        if max_seconds and time.time() - t_start > max_seconds:
            raise StopRun()
        return v

    functions.train_step = step
    functions.calc_psnr = psnr
    functions.get_optimizer = get_opt
    real_image = functions.wandb.Image
    functions.wandb.Image = lambda data=None, *a, **k: (real_image(data, *a, **k) if isinstance(data, np.ndarray) else None)

    work = workdir or tempfile.mkdtemp(prefix=f"gngf_main_{impl}_")
    os.makedirs(work, exist_ok=True)
    if not os.path.exists(os.path.join(work, "images")):
        os.symlink(os.path.join(ref_dir, "images"), os.path.join(work, "images"))
    cwd = os.getcwd()
    os.chdir(work)
    argv = sys.argv
    sys.argv = ["main.py", "-f", filename, "-s", str(param_id), "-e", str(param_id)]
    stopped = False
    try:
        runpy.run_path(os.path.join(ref_dir, "main.py"), run_name="__main__")
    except StopRun:
        stopped = True
    finally:
        sys.argv = argv
        os.chdir(cwd)
        functions.train_step, functions.calc_psnr, functions.get_optimizer = real_step, real_psnr, real_opt
        functions.wandb.Image = real_image
    info["stopped_by_time_budget"] = stopped
    info["epochs_run"] = len(rec["psnr"])
    info["wall_s"] = time.time() - t_start
    info["workdir"] = work
    saved = {}
    wdir = os.path.join(work, "weights")
    if os.path.isdir(wdir):
        for run_name in sorted(os.listdir(wdir)):
            for f in sorted(os.listdir(os.path.join(wdir, run_name))):
                sd = torch.load(os.path.join(wdir, run_name, f), map_location="cpu", weights_only=False)
                saved[f] = sorted(sd.keys()) if f != "whole_opt.pt" else sorted(sd.keys()) + \
                    [f"param_groups={len(sd['param_groups'])}", f"state={len(sd['state'])}"]
    info["saved_state_dicts"] = saved
    if impl == "dropin":
        import collision_handling_in_instantngp_b200 as pkg
        info["native_library"] = pkg.LIB_PATH
        info["gpu_launches"] = int(pkg.launch_count())
    if out:
        np.savez_compressed(out, info=json.dumps(info), **{k: np.asarray(v) for k, v in rec.items()}, **init)
    return info, rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["reference", "dropin"], required=True)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--param-id", type=int, default=4061)
    ap.add_argument("--filename", default="strawberry.jpeg")
    ap.add_argument("--ref-dir", default=os.path.join(HERE, "_ref"))
    ap.add_argument("--out", default="")
    ap.add_argument("--workdir", default="")
    ap.add_argument("--save-params", action="store_true", help="leave params.should_save_params on (functions.py:761-780)")
    ap.add_argument("--max-seconds", type=float, default=0.0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    info, rec = run(a.impl, a.epochs, a.out, os.path.abspath(a.ref_dir), a.device, a.param_id, a.filename,
                    a.save_params, a.workdir or None, a.max_seconds, a.threads, quiet=not a.verbose)
    info["final_psnr"] = float(rec["psnr"][-1]) if rec["psnr"] else None
    info["best_psnr"] = float(np.max(rec["psnr"])) if rec["psnr"] else None
    info["ms_per_epoch_median"] = float(np.median(rec["sec"]) * 1e3) if rec["sec"] else None
    print(json.dumps(info))


if __name__ == "__main__":
    main()
