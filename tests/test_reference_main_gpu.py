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


def run_main_concurrently(jobs):
    """jobs: [(impl, epochs, out, extra)] -> [(info, npz)]; the processes share the GPU (the reference's step is
    launch-bound, so three runs side by side take little longer than one)."""
    procs = []
    for impl, epochs, out, extra in jobs:
        cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_main.py"), "--impl", impl, "--epochs", str(epochs),
               "--out", out, *extra]
        env = dict(os.environ, WANDB_MODE="disabled", PYTHONUNBUFFERED="1")
        env.pop("PYTHONPATH", None)
        procs.append((out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)))
    res = []
    for out, p in procs:
        so, se = p.communicate(timeout=3000)
        assert p.returncode == 0, se[-4000:]
        res.append((json.loads(so.strip().splitlines()[-1]), np.load(out)))
    return res


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
    driven by the unmodified main.py for PSNR_EPOCHS epochs.

    The reference is NOT reproducible against itself: its scatter-adds are float atomics, the top-k selection is discrete,
    and two runs of the unmodified reference on the same B200 from the same seed agree to < 0.02 dB for ~80 epochs (PSNR
    7.2 -> 10 dB) and then drift apart -- up to 2.4 dB at the same epoch within 500 epochs (measured).
    So the 0.1 dB bar of north_star is decidable only on the deterministic prefix, and that is where it is asserted; after
    it the drop-in has to stay inside what the reference does to itself (window means, with the reference's own
    window-mean spread as the yardstick)."""
    _need_reference()
    budget = ["--max-seconds", str(REF_BUDGET_S)]
    (ia, a), (ib, b), (io, o) = run_main_concurrently([
        ("reference", PSNR_EPOCHS, str(tmp_path / "ref_a.npz"), budget),
        ("reference", PSNR_EPOCHS, str(tmp_path / "ref_b.npz"), budget),
        ("dropin", PSNR_EPOCHS, str(tmp_path / "ours.npz"), [])])
    n = min(len(a["psnr"]), len(b["psnr"]), len(o["psnr"]))
    assert n >= min(PSNR_EPOCHS, 100), (n, ia, ib)
    # same seed -> same pixel order and the same initial weights from both constructors (RNG consumption order)
    assert np.array_equal(a["shuffled_indices"], o["shuffled_indices"])
    for k in a.files:
        if k.startswith("init."):
            assert np.array_equal(a[k], o[k]), k
    pa, pb, po = a["psnr"][:n], b["psnr"][:n], o["psnr"][:n]
    spread = np.abs(pa - pb)
    prefix = int(np.argmax(spread > 0.02)) if (spread > 0.02).any() else n      # the reference agrees with itself
    diff_prefix = float(np.abs(po - pa)[:prefix].max()) if prefix else 0.0
    W = 100
    rows = []
    for s0 in range(0, n - W + 1, W):
        ma, mb, mo = pa[s0:s0 + W].mean(), pb[s0:s0 + W].mean(), po[s0:s0 + W].mean()
        rows.append((s0, float(ma), float(mb), float(mo)))
    win_spread = max([abs(r[1] - r[2]) for r in rows] + [0.0])
    marks = sorted(set(list(range(0, n, 50)) + [n - 1]))
    print(f"\nepochs compared: {n}; ms/epoch reference {np.median(a['sec']) * 1e3:.1f}, drop-in {np.median(o['sec']) * 1e3:.1f}")
    print("epoch      ", marks)
    print("reference A", np.round(pa[marks], 3))
    print("reference B", np.round(pb[marks], 3))
    print("drop-in    ", np.round(po[marks], 3))
    print(f"reference vs itself: identical to 0.02 dB for {prefix} epochs, then up to {spread.max():.3f} dB apart; "
          f"drop-in vs reference on that prefix: max {diff_prefix:.4f} dB")
    print("100-epoch window means (start, ref A, ref B, drop-in):", [tuple(round(v, 3) for v in r) for r in rows])
    report = {"epochs": n, "marks": marks, "ref_a": pa.tolist(), "ref_b": pb.tolist(), "ours": po.tolist(),
              "reference_self_agreement_epochs": prefix, "reference_spread_max_db": float(spread.max()),
              "dropin_vs_reference_on_prefix_max_db": diff_prefix, "window_means": rows,
              "ms_per_epoch": {"reference": float(np.median(a["sec"]) * 1e3), "dropin": float(np.median(o["sec"]) * 1e3)}}
    out_dir = os.path.join(ROOT, "gpurun_out")
    with open(os.path.join(out_dir if os.path.isdir(out_dir) else str(tmp_path), "psnr_parity.json"), "w") as f:
        json.dump(report, f)
    # (1) the bar, where it is decidable: 0.1 dB at EVERY epoch of the prefix on which the reference reproduces itself
    assert prefix >= 30, prefix
    assert diff_prefix <= 0.1, diff_prefix
    assert np.abs(o["mse"][:prefix] - a["mse"][:prefix]).max() <= 2e-3
    # (2) after it: window means within 0.1 dB + twice the reference's own window-mean spread of the nearer reference run.
    #     (Two trials on the B200 gave reference-vs-reference window spreads of 0.59 and 1.01 dB and drop-in distances of up
    #     to 0.86 dB to the nearer run; the yardstick is floored at 0.75 dB so that two reference runs that happen to stay
    #     close do not turn a chaotic trajectory into a failure.)
    yard = max(win_spread, 0.75)
    for s0, ma, mb, mo in rows:
        assert min(abs(mo - ma), abs(mo - mb)) <= 0.1 + 2.0 * yard, (s0, ma, mb, mo, win_spread)
    # ... and training got as far as the reference's
    assert po[-W:].max() >= min(pa[-W:].max(), pb[-W:].max()) - (0.1 + 2.0 * yard)


def _ref_step(impl, workload, points, steps, out, perturb=0.0):
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_step.py"), "--impl", impl, "--device", "cuda", "--points",
           str(points), "--steps", str(steps), "--warmup", "0", "--workload", json.dumps(workload), "--record", out,
           "--perturb", str(perturb)]
    env = dict(os.environ, WANDB_MODE="disabled")
    env.pop("PYTHONPATH", None)
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1200)
    assert p.returncode == 0, p.stderr[-4000:]
    return json.loads(p.stdout.strip().splitlines()[-1]), np.load(out)


def test_large_config_training_steps_match_the_reference(tmp_path):
    """The LARGE configuration in a training loop against the reference itself: BASELINE.json configs[3]'s model (16 levels
    16..8192, T = 2^14, K = 4, top-k-only probabilities) on 2 048 points of the 8192^2 lattice -- the drop-in takes its
    large-lattice route there (touched-node list over the 67 M-node box, skinny hidden layers, streaming tcgen05 forward /
    backward on fp16 planes) while the reference evaluates 131 072 rows x 16 384 slots densely (32 GB) -- for 12 steps of
    forward + the reference's own Loss + backward + the reference's own torch Adam, from the same seed.

    The trajectory is chaotic: the HPD sees integer coordinates up to 8 191, its logits are O(1e4), the softmax is one-hot
    and Adam's eps = 1e-15 (functions.py:104) turns gradients of 1e-12 into full-size steps.  How fast rounding-level
    noise grows is MEASURED on the reference itself: a second reference run starts from parameters multiplied by
    1 + 1e-7 N(0,1) (one fp32 ulp).  The drop-in must match the unperturbed reference to 1e-5 on step 0 (the same function
    of the same weights) and afterwards stay within three times the distance that one-ulp noise puts between the reference
    and itself."""
    _need_reference()
    sys.path.insert(0, ROOT)
    import bench
    w = {k: v for k, v in bench.WORKLOADS["cfg4_t14"].items()}
    steps, points = 12, 2048
    ir, zr = _ref_step("reference", w, points, steps, str(tmp_path / "ref.npz"))
    _, zp = _ref_step("reference", w, points, steps, str(tmp_path / "ref_p.npz"), perturb=1e-7)
    io, zo = _ref_step("dropin", w, points, steps, str(tmp_path / "ours.npz"))
    assert "baseline/_ref" in ir["models_file"].replace(os.sep, "/") and io["models_file"].endswith("models.py")
    for k in zr.files:
        if k.startswith("init."):
            assert np.array_equal(zr[k], zo[k]), k                       # same seed -> same initial weights
    lr_, lo_ = zr["losses"], zo["losses"]
    rel = np.abs(lo_ - lr_) / np.abs(lr_)
    noise = np.abs(zp["losses"] - lr_) / np.abs(lr_)
    print(f"\nlosses: reference {np.round(lr_, 5)}\n        drop-in   {np.round(lo_, 5)}\n        drop-in vs reference, per step:        "
          f"{np.array2string(rel, precision=1)}\n        reference vs itself + 1e-7 noise:      {np.array2string(noise, precision=1)}"
          f"\n        ms/step reference {ir['sec_per_step'] * 1e3:.0f}, drop-in {io['sec_per_step'] * 1e3:.0f}")
    assert rel[0] < 1e-5, rel
    envelope = 1e-5 + 3.0 * np.maximum.accumulate(noise)
    assert (rel <= envelope).all(), (rel, noise)


def test_full_run_fixture_of_the_reference():
    """tests/golden/ref_cuda_trajectory_4061.npz (baseline/full_run_fixture.py, recorded on one B200, all runs side by side):
    three runs of the UNMODIFIED reference (its CUDA-eager path; ~3 240 epochs each -- the 25-minute budget per process
    ended them before params.py's 5 000) and two complete runs of the drop-in (one early-stopped by the reference's own
    EarlyStopping at epoch 2 967, one through all 5 000 epochs), everything driven by the unmodified main.py from the same
    seed.  The statistic that survives the chaotic trajectory is the best PSNR of a run (what functions.py:761-780
    checkpoints), taken over the epochs every run covers: the drop-in's must lie within 0.1 dB + the reference's own
    run-to-run range of the reference's."""
    path = os.path.join(GOLDEN_DIR, "ref_cuda_trajectory_4061.npz")
    if not os.path.isfile(path):
        pytest.skip("no full-run fixture recorded yet")
    z = np.load(path)
    refs = [z[f"ref_psnr_{i}"] for i in range(int(z["ref_runs"]))]
    ours = [z[f"ours_psnr_{i}"] for i in range(int(z["ours_runs"]))]
    n = min(len(r) for r in refs)
    ref_best = np.array([r[:n].max() for r in refs])
    our_best = np.array([o[:n].max() for o in ours])
    rng = float(ref_best.max() - ref_best.min())
    print(f"\nbest PSNR over the first {n} epochs: reference {np.round(ref_best, 3)} (range {rng:.3f} dB), drop-in "
          f"{np.round(our_best, 3)}; drop-in run lengths {[len(o) for o in ours]}, best over a complete run "
          f"{max(float(o.max()) for o in ours):.3f} dB (reference README.md:30: 20.331 dB)")
    assert len(ref_best) >= 2 and len(our_best) >= 1 and n >= 3000
    for v in our_best:
        assert ref_best.min() - 0.1 - rng <= v <= ref_best.max() + 0.1 + rng, (v, ref_best)
    # and at fixed epochs the drop-in's 200-epoch means stay within the band of the reference runs (+- its width)
    for e in (1000, 2000, 3000):
        rm = np.array([r[e - 200:e].mean() for r in refs])
        w = float(rm.max() - rm.min())
        for o in ours:
            if len(o) >= e:
                assert rm.min() - 0.1 - w <= o[e - 200:e].mean() <= rm.max() + 0.1 + w, (e, rm, float(o[e - 200:e].mean()))
