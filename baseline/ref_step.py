"""BENCH / TEST INFRASTRUCTURE ONLY -- times the UNMODIFIED reference's hot path (baseline/_ref, see fetch_ref.py) on
synthetic batches of a bench.py workload, on the host cores or on cuda:0.

    python baseline/ref_step.py --device cpu  --points 256 --steps 3 --warmup 1 --workload '{"L": 16, ...}'
    python baseline/ref_step.py --device cuda --points 2048 ...

One step = the body of the reference's batch loop (functions.py:203-281) with the reference's own objects:
`GeneralNeuralGaugeFields(...)` (models.py:239-484), `Loss(delta=1, gamma, epsilon)` (utils.py:78-174), the loss
assembly of functions.py:243-245, `loss.backward()`, `get_optimizer(...).step()` (functions.py:96-127, 281).  The
epoch-level bookkeeping of train_step (collision counting, image assembly) is not part of the metric ("fwd+bwd GNGF
hash encode + MLP") and is not timed.  Always a separate process: functions.py forces the default device and seeds
the global generator at import time.  Prints one JSON line.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from run_main import install_matplotlib_stub  # noqa: E402


def load_reference(ref_dir, device, impl="reference"):
    """impl = "reference": everything is the reference's.  impl = "dropin": `models` resolves to this repository's drop-in
    module (repository root ahead of the reference on sys.path, as in run_main.py); functions / utils / params stay the
    reference's, so Loss and get_optimizer are its own."""
    import torch
    os.environ["WANDB_MODE"] = "disabled"
    install_matplotlib_stub()
    for name in ("functions", "models", "utils", "params"):
        sys.modules.pop(name, None)
    root = os.path.dirname(HERE)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != root]       # never this repository's models.py ...
    sys.path.insert(0, ref_dir)
    if impl == "dropin":
        sys.path.insert(0, root)                                                    # ... unless it is what is asked for
    real = torch.set_default_device
    try:
        if device == "cpu":
            torch.set_default_device = lambda *_a, **_k: None
        functions = importlib.import_module("functions")
        utils = importlib.import_module("utils")
        models = importlib.import_module("models")
    finally:
        torch.set_default_device = real
    dev = torch.device(device)
    for m in (functions, models, utils):
        m.device = dev
    if impl == "dropin":
        assert not os.path.realpath(models.__file__).startswith(os.path.realpath(ref_dir)), models.__file__
    else:
        assert os.path.realpath(models.__file__).startswith(os.path.realpath(ref_dir)), models.__file__
    assert os.path.realpath(utils.__file__).startswith(os.path.realpath(ref_dir)), utils.__file__
    return functions, models, utils


def make_step(functions, models, utils, w, x, y, device):
    """(step closure, net) for workload dict `w` (bench.py WORKLOADS entry) on tensors x (P,2), y (P,3)."""
    import torch
    torch.manual_seed(65535)
    net = models.GeneralNeuralGaugeFields(
        input_dim=2, hash_table_size=w["T"], num_levels=w["L"], n_min=w["n_min"], n_max=w["n_max"],
        MLP_hidden_layers_widths=list(w["mlp"]), HPD_hidden_layers_widths=list(w["hpd"]), HPD_out_features=w["T"],
        feature_dim=w["F"], topk_k=w["K"], should_keep_topk_only=bool(w["topk_only"]))
    if device == "cpu":
        net = net.to("cpu")
    loss_fn = utils.Loss(delta=1, gamma=w["gamma"], epsilon=w["epsilon"])
    opt = functions.get_optimizer(net=net, encoding_lr=w["lr"]["encoding"], HPD_lr=w["lr"]["hpd"], MLP_lr=w["lr"]["mlp"],
                                  encoding_weight_decay=w["wd"]["encoding"], HPD_weight_decay=w["wd"]["hpd"],
                                  MLP_weight_decay=w["wd"]["mlp"])
    empty = torch.tensor([], device=device)
    net.train()

    def step():
        opt.zero_grad()
        output, probs, indices_topk, _ = net(x, 1.0, should_calc_counts=False)
        mse, kl, coll = loss_fn(output, y, probs.shape[-1], probs, empty, empty)
        loss = w["l_mse"] * mse                                                   # functions.py:243-245
        loss = loss + ((w["l_js_kl"] * kl) + (w.get("l_collisions", 1e-3) * coll if coll.nelement() != 0 else 1)).sum(0)
        loss.backward()
        opt.step()
        return loss

    return step, net


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--points", type=int, required=True)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--workload", required=True, help="JSON of a bench.py WORKLOADS entry")
    ap.add_argument("--ref-dir", default=os.path.join(HERE, "_ref"))
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--impl", default="reference", choices=["reference", "dropin"])
    ap.add_argument("--record", default="", help="npz: the loss of every step, the initial and the final parameters")
    ap.add_argument("--perturb", type=float, default=0.0, help="multiply every parameter by 1 + PERTURB * N(0,1) before the "
                    "first step (how fast does rounding-level noise grow in this configuration?)")
    a = ap.parse_args()
    import torch
    w = json.loads(a.workload)
    if a.device == "cpu":
        torch.set_num_threads(a.threads or os.cpu_count() or 1)
    functions, models, utils = load_reference(os.path.abspath(a.ref_dir), a.device, a.impl)
    sys.path.insert(0, os.path.dirname(HERE))
    from bench import make_inputs                                               # the same synthetic batch as the GPU arm
    x_np, y_np = make_inputs(dict(w, P=a.points), 65535)
    x = torch.from_numpy(x_np).to(a.device)
    y = torch.from_numpy(y_np).to(a.device)
    step, net = make_step(functions, models, utils, w, x, y, a.device)
    if a.perturb:
        g = torch.Generator(device=a.device).manual_seed(12345)
        with torch.no_grad():
            for prm in net.parameters():
                prm.mul_(1.0 + a.perturb * torch.randn(prm.shape, generator=g, device=prm.device))
    init = {"init." + k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
    losses = []

    def sync():
        if a.device == "cuda":
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        losses.append(float(step().detach()))
    sync()
    times = []
    for _ in range(a.steps):
        sync()
        t0 = time.perf_counter()
        loss = step()
        losses.append(float(loss.detach()))
        sync()
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    out = {"device": a.device, "points": a.points, "steps": a.steps, "warmup": a.warmup, "sec_per_step": sec,
           "samples_per_s": a.points / sec, "threads": int(torch.get_num_threads()) if a.device == "cpu" else None,
           "loss": float(loss.detach()), "models_file": os.path.realpath(models.__file__)}
    out["impl"] = a.impl
    if a.record:
        final = {"final." + k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
        np.savez_compressed(a.record, losses=np.asarray(losses), **init, **final)
    if a.device == "cuda":
        out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
        out["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
