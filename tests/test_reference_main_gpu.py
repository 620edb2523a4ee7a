"""GPU: the reference's UNMODIFIED main.py / functions.py / utils.py / params.py (baseline/_ref, see
baseline/fetch_ref.py) driving this repository's drop-in `models` module, and PSNR parity against the reference's
own CUDA-eager path on the same GPU, seed, image and parameter ID (BASELINE.json north_star; reference README.md:30:
strawberry.jpeg, grid-search ID 4061).

Every run is a separate process (`baseline/run_main.py`): main.py star-imports `models`, forces the default device
and seeds the global generator at import time (functions.py:43-52), none of which can be undone inside one process.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from golden_util import GOLDEN_DIR

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
PSNR_EPOCHS = int(os.environ.get("GNGF_PSNR_EPOCHS", "500"))
REF_BUDGET_S = float(os.environ.get("GNGF_PSNR_REF_BUDGET_S", "200"))


def _need_reference():
    if not os.path.isfile(os.path.join(REF, "main.py")):
        pytest.skip("baseline/_ref is absent (python baseline/fetch_ref.py copies it where /root/reference exists)")


def run_main(impl, epochs, out, extra=()):
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_main.py"), "--impl", impl, "--epochs", str(epochs),
           "--out", out, *extra]
    env = dict(os.environ, WANDB_MODE="disabled", PYTHONUNBUFFERED="1")
    env.pop("PYTHONPATH", None)
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=3000)
    assert p.returncode == 0, p.stderr[-4000:]
    info = json.loads(p.stdout.strip().splitlines()[-1])
    return info, np.load(out)


def test_unmodified_main_runs_on_the_dropin_module(tmp_path):
    """main.py:3,72-74 -> functions.py:540-556 (ctor), 203 (forward), 227-245 (Loss on the returned probs), 272-281
    (backward + stock torch Adam), 327 (collisions on the float32 torch.empty buffer of functions.py:179,216),
    761-780 (five state-dict saves), epoch 0 with should_calc_counts (histograms, functions.py:670)."""
    _need_reference()
    info, z = run_main("dropin", 3, str(tmp_path / "ours.npz"), ["--save-params", "--workdir", str(tmp_path / "w")])
    assert info["epochs_run"] == 3 and not info["stopped_by_time_budget"]
    assert info["net_class"] == "collision_handling_in_instantngp_b200.models.GeneralNeuralGaugeFields"
    assert os.path.realpath(info["models_file"]) == os.path.realpath(os.path.join(ROOT, "models.py"))
    assert info["optimizer"].startswith("torch.optim")                     # the reference's own optimizer, untouched
    assert info["native_library"].endswith("libgngf_sm100.so") and info["gpu_launches"] > 0
    # train_step hands calc_hash_collisions its float32 torch.empty buffer (w*h, L, 4, K / batch_size)
    assert info["collisions_input"] == {"dtype": "torch.float32", "shape": [172212, 4, 4, 12], "device": "cuda:0"}
    # the five files of functions.py:767-780, with the reference's own state-dict keys (recorded from the unmodified
    # reference by `run_main.py --impl reference --save-params`, tests/golden/ref_state_dict_keys.json)
    want = json.load(open(os.path.join(GOLDEN_DIR, "ref_state_dict_keys.json")))
    assert info["saved_state_dicts"] == want
    assert np.isfinite(z["psnr"]).all() and np.isfinite(z["loss"]).all()
    assert z["collisions"].shape == (3, 4)


def test_psnr_trajectory_matches_the_reference_cuda_eager_path(tmp_path):
    """Same GPU, seed (functions.py:43-47), image, parameter ID: the reference's own ATen CUDA path vs the drop-in, both
    driven by the unmodified main.py for PSNR_EPOCHS epochs.  The reference's scatter-adds are atomics (non-deterministic
    order), so its own run-to-run spread is measured first (two reference runs); the bar is 0.1 dB (north_star)."""
    _need_reference()
    budget = ["--max-seconds", str(REF_BUDGET_S)]
    ia, a = run_main("reference", PSNR_EPOCHS, str(tmp_path / "ref_a.npz"), budget)
    ib, b = run_main("reference", PSNR_EPOCHS, str(tmp_path / "ref_b.npz"), budget)
    io, o = run_main("dropin", PSNR_EPOCHS, str(tmp_path / "ours.npz"))
    n = min(len(a["psnr"]), len(b["psnr"]), len(o["psnr"]))
    assert n >= min(PSNR_EPOCHS, 100), (n, ia, ib)
    # same seed -> same pixel order and the same initial weights from both constructors (RNG consumption order)
    assert np.array_equal(a["shuffled_indices"], o["shuffled_indices"])
    for k in a.files:
        if k.startswith("init."):
            assert np.array_equal(a[k], o[k]), k
    pa, pb, po = a["psnr"][:n], b["psnr"][:n], o["psnr"][:n]
    spread = np.abs(pa - pb)
    diff = np.minimum(np.abs(po - pa), np.abs(po - pb))
    marks = sorted(set(list(range(0, n, 50)) + [n - 1]))
    print(f"\nepochs compared: {n}; ms/epoch reference {np.median(a['sec']) * 1e3:.1f}, drop-in {np.median(o['sec']) * 1e3:.1f}")
    print("epoch      ", marks)
    print("reference A", np.round(pa[marks], 3))
    print("reference B", np.round(pb[marks], 3))
    print("drop-in    ", np.round(po[marks], 3))
    print(f"reference run-to-run spread: max {spread.max():.4f} dB; drop-in vs nearest reference run: max {diff.max():.4f} dB "
          f"(at the marks: {diff[marks].max():.4f})")
    with open(os.path.join(ROOT, "gpurun_out", "psnr_parity.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out"))
              else str(tmp_path / "psnr_parity.json"), "w") as f:
        json.dump({"epochs": n, "marks": marks, "ref_a": pa.tolist(), "ref_b": pb.tolist(), "ours": po.tolist(),
                   "spread_max": float(spread.max()), "diff_max": float(diff.max()),
                   "ms_per_epoch": {"reference": float(np.median(a["sec"]) * 1e3), "dropin": float(np.median(o["sec"]) * 1e3)}}, f)
    # the bar: within 0.1 dB of the reference at every epoch (plus whatever the reference differs from itself)
    assert (diff <= 0.1 + spread).all(), (float(diff.max()), float(spread.max()))
    assert abs(po[n - 1] - pa[n - 1]) <= 0.1 + spread.max()
    # the MSE half of the loss tracks as well
    assert np.abs(o["mse"][:n] - a["mse"][:n]).max() <= 2e-3 + np.abs(a["mse"][:n] - b["mse"][:n]).max()


def test_full_run_fixture_of_the_reference():
    """tests/golden/ref_cuda_trajectory_4061.npz: one complete (early-stopped or 5 000-epoch) run of the unmodified
    reference on a B200 next to a complete run of the drop-in, recorded by the builder with baseline/run_main.py.
    Checked here: the recorded final / best PSNR of the two agree within 0.1 dB + the reference's recorded spread."""
    path = os.path.join(GOLDEN_DIR, "ref_cuda_trajectory_4061.npz")
    if not os.path.isfile(path):
        pytest.skip("no full-run fixture recorded yet")
    z = np.load(path)
    ref, ours = z["ref_psnr"], z["ours_psnr"]
    spread = float(z["ref_spread_db"])
    print(f"\nfull run: reference {len(ref)} epochs, best {ref.max():.3f} dB, final {ref[-1]:.3f} dB; "
          f"drop-in {len(ours)} epochs, best {ours.max():.3f} dB, final {ours[-1]:.3f} dB; reference spread {spread:.3f} dB")
    assert abs(ref.max() - ours.max()) <= 0.1 + spread
