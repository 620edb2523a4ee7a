#!/usr/bin/env python
"""bench.py -- train samples/s of the GNGF hot path (forward + loss + backward + Adam) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4_t14]

One "step" = one pass of the hot path over one batch of synthetic coordinates:
GeneralNeuralGaugeFields.forward (models.py:394-484) -> the reference's loss assembly (utils.py:78-174,
functions.py:243-245) -> backward -> Adam (functions.py:96-127).

Default workload = BASELINE.json configs[3], the configuration the metric's "at 1/2/4/8 B200" names: a synthetic
8192 x 8192 pixel lattice, 2^22 points per step, 16 levels x 2 features, K = 4, HPD 2-32-64-128-T, decoder 32-64-64-3,
top-k-only probabilities -- with the table size T = 2^14 stated in `config` ("cfg4_t14"): at the survey's T = 2^19 one
step executes 6e16 tensor-core FLOP (~50 s on one B200), which does not fit a default bench run; that line is
builder-run (`--workload cfg4`, profiles/).  The step's 2^22 points are SPLIT over the N ranks (strong scaling): every
rank runs the point passes on its 2^22 / N points and 1 / N of the lattice nodes any rank touches through the HPD
(dp.NodeSharding).  At N = 1 the line also carries `secondary`: configs[1] (cfg2, strawberry / ID 4061 shapes) and
configs[2] (cfg3: macaw shape, T = 2^19 as specified) measured in the same process.

JSON line (rank 0): `value` is device time with inputs resident in HBM (CUDA events per step, L2 flushed between
steps); `e2e` the same step through the module API with HOST (pinned) inputs copied in and the loss read back every
step; `roofline` describes the dominant kernel of the step (per-call CUDA events, algorithmic cost of DESIGN.md
section 4); `cpu_baseline` = the UNMODIFIED reference (baseline/_ref: its GeneralNeuralGaugeFields + Loss + Adam) on
this box's host cores on a bounded sample of the workload; `gpu_eager_reference` = the same unmodified reference on
cuda:0 (its ATen eager path) at the largest sample that fits -- the number the drop-in replaces.  `--impl reference`
times the reference's CPU path alone.  When baseline/_ref is absent the reference legs fall back to
oracle/torch_port.py (kind "port").
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec (fwd+bwd GNGF hash encode+MLP)"
_COMMON = dict(K=4, F=2, hpd=[32, 64, 128], mlp=[64, 64], gamma=-2.0, epsilon=1.0, l_mse=1.0, l_js_kl=1.0,
               lr=dict(encoding=1e-4, hpd=1e-3, mlp=1e-3), wd=dict(encoding=0.0, hpd=1e-6, mlp=1e-6))
WORKLOADS = {
    # BASELINE.json configs[1]: strawberry.jpeg + param ID 4061 (README.md:15-18, params.py:26-51); one batch = a third of
    # the image (batch_size = 1/3).  Per-rank batch fixed as N grows (weak scaling: the published shapes)
    "cfg2": dict(_COMMON, P=57404, L=4, n_min=8, n_max=32, T=256, lattice_hw=(508, 339), topk_only=False,
                 cpu_sample=57404, scaling="weak", names="configs[1]: strawberry.jpeg shapes, param ID 4061"),
    # BASELINE.json configs[2]: macaw.jpg (508x339 = 172 212 px < 2^18: the whole image is one batch), 16 levels,
    # table size 2^19, top-k-only probabilities (the full distribution would be 134 MB per sample)
    "cfg3": dict(_COMMON, P=172212, L=16, n_min=16, n_max=508, T=2 ** 19, lattice_hw=(508, 339), topk_only=True,
                 cpu_sample=4, scaling="strong", names="configs[2]: macaw.jpg shape, 16 levels, T = 2^19"),
    # same image shape with a mid-size table (fits the dense path as well; used to compare the two HPD paths)
    "cfg3_t14": dict(_COMMON, P=172212, L=16, n_min=16, n_max=508, T=2 ** 14, lattice_hw=(508, 339), topk_only=True,
                     cpu_sample=64, scaling="strong", names="configs[2] shape with T = 2^14"),
    # BASELINE.json configs[3]: synthetic 8192 x 8192 lattice, 2^22 points per step, 16 levels x 2 features
    "cfg4_t14": dict(_COMMON, P=2 ** 22, L=16, n_min=16, n_max=8192, T=2 ** 14, lattice_hw=(8192, 8192), topk_only=True,
                     cpu_sample=128, scaling="strong",
                     names="configs[3]: synthetic 8192x8192 lattice, 2^22 points/step, 16 levels x 2 features, T = 2^14"),
    "cfg4": dict(_COMMON, P=2 ** 22, L=16, n_min=16, n_max=8192, T=2 ** 19, lattice_hw=(8192, 8192), topk_only=True,
                 cpu_sample=4, scaling="strong",
                 names="configs[3]: synthetic 8192x8192 lattice, 2^22 points/step, 16 levels x 2 features, T = 2^19"),
}
DEFAULT_WORKLOAD = "cfg4_t14"


def make_inputs(w, seed, rank=0):
    """Synthetic batch of the workload's shape: coordinates are pixel-lattice points (row, col)/(max(h,w)-1) as in
    main.py:42-51, drawn without replacement in a seeded random order; targets uniform in [0,1)."""
    h, wd = w["lattice_hw"]
    rng = np.random.default_rng(seed + 7919 * rank)
    flat = rng.permutation(h * wd)[: w["P"]]
    x = np.stack([flat // wd, flat % wd], 1).astype(np.float32) / np.float32(max(h, wd) - 1)
    y = rng.random((w["P"], 3), dtype=np.float32)
    return x, y


def static_config(name, w, world):
    """The workload as both arms state it (identical dicts: the driver compares them)."""
    return {"workload": name, "names": w["names"], "points_per_step": w["P"] * (world if w["scaling"] == "weak" else 1),
            "levels": w["L"], "table_size": w["T"], "topk_k": w["K"], "feature_dim": w["F"], "hpd": [2, *w["hpd"], w["T"]],
            "decoder": [w["L"] * w["F"], *w["mlp"], 3], "n_min": w["n_min"], "n_max": w["n_max"],
            "step": "forward+loss+backward+Adam", "parallelism": f"dp{world}"}


# ------------------------------------------------------------------------------------------------------------
# reference legs: the UNMODIFIED reference (baseline/_ref) in a child process; oracle port when it is absent
# ------------------------------------------------------------------------------------------------------------
def reference_available():
    return os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "models.py"))


def time_reference(w, device, points, steps, warmup, timeout=1500):
    """samples/s record of baseline/ref_step.py (the reference's own classes; see its docstring), or None."""
    keep = {k: v for k, v in w.items() if not k.startswith("_")}
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_step.py"), "--device", device, "--points", str(points),
           "--steps", str(steps), "--warmup", str(warmup), "--workload", json.dumps(keep)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "PYTHONPATH"):
        env.pop(k, None)
    if device == "cuda":
        env["CUDA_VISIBLE_DEVICES"] = os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {timeout} s"}
    if p.returncode != 0:
        return {"error": p.stderr.strip().splitlines()[-1][:300] if p.stderr.strip() else f"rc {p.returncode}"}
    return json.loads(p.stdout.strip().splitlines()[-1])


def time_cpu_port(w, sample_P, seed, steps, warmup):
    """Fallback when baseline/_ref is absent: seconds per step of oracle/torch_port.py (the reference's ATen ops +
    autograd restated on PyTorch-CPU, all host threads) on `sample_P` coordinates of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import gngf_oracle as O
    import torch_port as TP
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = make_inputs(dict(w, P=sample_P), seed)
    step = TP.make_step(w, x, y, O.level_resolutions(w["n_min"], w["n_max"], w["L"]), seed)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def cpu_reference_record(w, steps, warmup, sample_P=None):
    """cpu_baseline object: the reference's CPU path on a bounded sample, all host threads."""
    sample_P = int(sample_P or w["cpu_sample"])
    what = f"{sample_P} of the workload's {w['P']} coordinates per step, {steps} steps after {warmup} warm-up, forward+loss+backward+Adam"
    if reference_available():
        r = time_reference(w, "cpu", sample_P, steps, warmup)
        if r and "error" not in r:
            return {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "kind": "reference",
                    "sample": what + "; the UNMODIFIED reference (baseline/_ref: GeneralNeuralGaugeFields + Loss + "
                                     "get_optimizer, PyTorch-CPU, all host threads)", "ms_per_step": r["sec_per_step"] * 1e3}
        err = r["error"] if r else "no output"
    else:
        err = "baseline/_ref absent"
    import torch
    sec = time_cpu_port(w, sample_P, 65535, steps, warmup)
    return {"value": sample_P / sec, "unit": "samples/s", "cores": int(torch.get_num_threads()), "kind": "port",
            "sample": what + f"; oracle/torch_port.py (PyTorch-CPU restatement; reference unavailable: {err})",
            "ms_per_step": sec * 1e3}


def gpu_reference_points(w):
    """Largest power-of-two sample whose (rows, T) fp32 tensors -- ~4 live copies at the peak of the reference's forward
    + autograd (measured: 16.1 GB at 1 024 points of configs[3] with T = 2^14), 5 budgeted -- stay below ~80 GB."""
    per_point = 4 * w["L"] * w["T"] * 4 * 5
    p = max(1, int(80e9 // per_point))
    return int(min(w["P"], 2 ** int(np.log2(p)))) if p < w["P"] else w["P"]


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled through NVML in a background thread during the timed regions.
    (An `nvidia-smi -lms` child process was measured to slow every CUDA launch of the timed process by ~5x
    through driver-lock contention; in-process NVML queries at 25 ms do not.)"""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.power = index, [], set(), []
        self.max_sm, self.h, self._stop, self.err = None, None, threading.Event(), None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            self._stop.wait(0.025)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


class CallProfiler:
    """Times every C-ABI call with CUDA events on the launching stream (installed as _lib.PROFILER)."""

    def __init__(self, torch):
        self.torch, self.events = torch, []

    def record(self, name, fn, args):
        a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        self.events.append((name, args, a, b))
        return rc

    def summary(self):
        self.torch.cuda.synchronize()
        agg = {}
        for name, args, a, b in self.events:
            key = (name, cost_key(name, args))
            t, n = agg.get(key, (0.0, 0))
            agg[key] = (t + a.elapsed_time(b), n + 1)
        return agg


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` captures (profiles/r01_*_ncu_full.txt); None where no capture exists yet
NCU_TRAFFIC = {("cfg2", "gngf_mlp3_bwd"): 2606080, ("cfg2", "gngf_mlp3_fwd"): 1892096,
               ("cfg2", "gngf_mlp3_tc_bwd"): 4250112, ("cfg2", "gngf_mlp3_tc_fwd"): 1924608,
               # configs[3] / T = 2^14 (profiles/r02_cfg4_t14_stream_skip_ncu.txt): the dh pass of the streaming backward
               # (31.48 GB read + 15.32 GB written; the dW3 pass of the same C call moves the same planes again: ncu
               # returned no counters for it), the streaming forward pass (15.43 + 8.62 GB)
               ("cfg4_t14", "gngf_hpd_stream_bwd_nodes"): 46799761000,
               ("cfg4_t14", "gngf_hpd_stream_fwd_refined"): 24050862000,
               ("cfg3_t14", "gngf_hpd_stream_bwd"): 441328128,
               ("cfg3_t14", "gngf_hpd_stream_fwd"): 154538000, ("cfg3_t14", "gngf_tc_gemm_bf16x3"): 2802181000}


def cost_key(name, args):
    """Shape signature of a call -> used to attach algorithmic bytes / FLOPs."""
    if name == "gngf_linear_fwd":
        return tuple(int(v) for v in args[3:6])
    if name == "gngf_linear_bwd":
        return tuple(int(v) for v in args[3:6]) + (args[7] is not None,)
    if name == "gngf_tc_gemm_bf16x3":
        return tuple(int(v) for v in args[3:6])
    if name == "gngf_peer_allreduce":
        return (int(args[3]), int(args[6]))                  # (world, floats)
    if name in ("gngf_hpd_stream_fwd", "gngf_hpd_stream_fwd_refined"):
        return (_rows_arg(name, args),)
    if name in ("gngf_hpd_stream_bwd", "gngf_hpd_stream_bwd_nodes"):
        return (_rows_arg(name, args),)
    return ()


def _rows_arg(name, args):
    """Rows (lattice nodes) a streaming HPD call processes: under node parallelism a rank's share of the node list."""
    if name == "gngf_hpd_stream_fwd":
        return int(args[3])
    if name == "gngf_hpd_stream_fwd_refined":
        return int(args[7])
    if name == "gngf_hpd_stream_bwd":
        return int(args[8])
    return int(args[9])


def algorithmic_cost(name, key, w, lat):
    """(bound, amount per launch, unit): ALGORITHMIC bytes (hbm) or FLOPs (tensor) of one launch."""
    P, L, F, K, T = w["P"], w["L"], w["F"], w["K"], w["T"]
    U, S = lat.num_nodes, lat.num_level_nodes
    Ua = w.get("_active_nodes") or U      # rows of the HPD chain (ops.active_nodes: the nodes the batch touches)
    if name.startswith("gngf_hpd_stream_") and key:
        Ua = key[0]                       # the rows this call processed (node parallelism: 1/world of the list)
    if name == "gngf_peer_allreduce":
        # every rank reads the other ranks' slices over NVLink: (world - 1) * n * 4 bytes against 900 GB/s per direction
        return "nvlink", float((key[0] - 1) * key[1] * 4)
    if name in ("gngf_mlp3_tc_fwd", "gngf_mlp3_tc_bwd"):
        # EXECUTED tensor-core FLOPs: inputs padded to 16, output layer padded to 16 columns;
        # forward: 6 split products; backward: 3 split products over recompute (2 layers) + dA2, dA1, dX + dW2, dW1, dW0
        inp = (L * F + 15) // 16 * 16
        if name == "gngf_mlp3_tc_fwd":
            return "tensor", 6 * 2.0 * P * (inp * 64 + 64 * 64 + 64 * 16)
        return "tensor", 3 * 2.0 * P * ((inp * 64 + 64 * 64) + (16 * 64 + 64 * 64 + 64 * inp) + (64 * 16 + 64 * 64 + 64 * inp))
    if name in ("gngf_mlp3_fwd", "gngf_mlp3_bwd"):
        dims = [L * F, *w["mlp"], 3]
        flops = 2.0 * P * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        return "tensor", flops * (1 if name == "gngf_mlp3_fwd" else 3)   # bwd: recompute + dX + dW
    if name == "gngf_linear_fwd":
        M, N, Kd = key
        return "tensor", 2.0 * M * N * Kd
    if name == "gngf_linear_bwd":
        M, N, Kd, has_dx = key
        return "tensor", 2.0 * M * N * Kd * (2 if has_dx else 1)
    if name.startswith("gngf_hpd_small_"):
        # fused small-lattice HPD (fp32 CUDA cores; GEMM-shaped work, so it is put against the tensor peak): the layer
        # chain on U nodes, forward; the dX chain in the backward (its dW products are gngf_linear_bwd launches)
        dims = [2, *w["hpd"], T]
        return "tensor", 2.0 * U * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    kd = w["hpd"][-1]
    # tensor-core kernels: EXECUTED FLOPs = 6 split-precision passes over the useful 2*M*N*K;
    # the useful figure is reported next to it (roofline.useful_tflops)
    if name == "gngf_hpd_stream_fwd":
        return "tensor", 6 * 2.0 * Ua * T * kd
    if name == "gngf_hpd_stream_fwd_refined":      # two planes, three split products (+ an fp32 refinement of 8 candidates)
        return "tensor", 3 * 2.0 * Ua * T * kd
    if name == "gngf_tc_gemm_bf16x3":
        return "tensor", 6 * 2.0 * float(key[0]) * key[1] * key[2]
    if name in ("gngf_hpd_stream_bwd", "gngf_hpd_stream_bwd_nodes"):
        # two fused passes (dh, dW3), each: one screening product of the logits on every tile, the other two on the
        # tiles it cannot rule out, the second product (3 split products) on the tiles whose E is not all zero --
        # products per tile counted on the device during the profiled steps (gngf_hpd_stream_bwd_stats; 6 if nothing
        # is skipped); useful work = the two gradient products, 2 * 2*U*T*kd
        return "tensor", 2 * w.get("_stream_products_per_tile", 6.0) * 2.0 * Ua * T * kd
    table = {
        # per point: x (8) + per level 4 node-feature gathers (4*F*4) + enc row (F*4) + 4 multiplicity atomics (4*4)
        "gngf_encode_fwd": P * (8 + L * (4 * F * 4 + F * 4 + 16)),
        # per point: x (8) + per level d enc (F*4) + 4 vector reductions (4*F*4)
        "gngf_encode_bwd": P * (8 + L * (F * 4 + 4 * F * 4)),
        # API output idx_topk (P,L,4,K) int64 written, (K int32 read per row)
        "gngf_lattice_gather_rows_i64": P * L * 4 * K * (8 + 4) + P * 8,
        "gngf_lattice_gather_rows": P * L * 4 * K * (4 + 4) + P * 8,
        "gngf_node_features_fwd": S * (K * (8 + F * 4) + F * 4),
        "gngf_node_features_bwd": S * (F * 4 + K * (8 + 2 * F * 4 + 4)),
        "gngf_softmax_topk_fwd": U * (2 * T * 4 + K * 8),
        "gngf_hpd_dlogits": U * (2 * T * 4 + K * 12) + L * T * 4,
        "gngf_lattice_colsum": S * 4 + U * T * 4 + L * T * 4,
        "gngf_sigmoid_bwd": P * 3 * 4 * 3,
        # rgb + target read, d_rgb written; column sums read, their adjoint written
        "gngf_loss_fwd_bwd": P * 3 * 12 + L * (K if w["topk_only"] else T) * 8,
        "gngf_loss_parts": P * 3 * 12 + L * (K if w["topk_only"] else T) * 8,
        "gngf_hpd_first_layer_fwd_nodes": Ua * w["hpd"][0] * 4,
        "gngf_hpd_first_layer_bwd_nodes": Ua * w["hpd"][0] * 4,
        # P*L*4 bit sets on a U-bit map; the compaction reads the map three times and writes the ids
        "gngf_lattice_mark_nodes": P * 8 + U / 8,
        "gngf_compact_nodes": 2 * U / 8 + Ua * 4,
        "gngf_scatter_node_rows": Ua * (4 + K * 8),
        "gngf_gather_node_adjoints": Ua * (4 + K * 8) + Ua * L * 4,
        "gngf_bitmap_or": U / 8 * 3,
        "gngf_cell_to_node_counts": S * 4 * 5,
        "gngf_split_bf16x3": 0.0,
        # |x| max pass (read) + split pass (read, two fp16 planes written)
        "gngf_split_f16x2": 0.0,
    }
    return "hbm", float(table.get(name, 0))


def useful_tflops(name, achieved, w):
    """The share of the executed tensor-core FLOPs that the fp32 algorithm asks for (one pass per product)."""
    if name in ("gngf_hpd_stream_fwd", "gngf_tc_gemm_bf16x3"):
        return achieved / 6
    if name == "gngf_hpd_stream_fwd_refined":
        return achieved / 3
    if name in ("gngf_hpd_stream_bwd", "gngf_hpd_stream_bwd_nodes"):
        return achieved / w.get("_stream_products_per_tile", 6.0)   # executed products per pass for 1 gradient product
    if name in ("gngf_mlp3_tc_fwd", "gngf_mlp3_tc_bwd"):
        dims = [w["L"] * w["F"], *w["mlp"], 3]
        useful = 2.0 * sum(a * b for a, b in zip(dims[:-1], dims[1:])) * (1 if name == "gngf_mlp3_tc_fwd" else 2)
        inp = (dims[0] + 15) // 16 * 16
        if name == "gngf_mlp3_tc_fwd":
            executed = 6 * 2.0 * (inp * 64 + 64 * 64 + 64 * 16)
        else:
            executed = 3 * 2.0 * ((inp * 64 + 64 * 64) + (16 * 64 + 64 * 64 + 64 * inp) + (64 * 16 + 64 * 64 + 64 * inp))
        return achieved * useful / executed
    return None


class Runner:
    """One workload on this rank: model, optimizer, inputs and the step closure."""

    def __init__(self, torch, dist, name, w, rank, world, dev):
        from collision_handling_in_instantngp_b200 import dp
        from collision_handling_in_instantngp_b200.loss import fused_total_loss
        from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
        from collision_handling_in_instantngp_b200.optim import FusedAdam
        self.torch, self.dist, self.dp = torch, dist, dp
        self.name, self.w, self.rank, self.world, self.dev = name, w, rank, world, dev
        torch.manual_seed(65535)
        self.net = net = GeneralNeuralGaugeFields(
            input_dim=2, hash_table_size=w["T"], num_levels=w["L"], n_min=w["n_min"], n_max=w["n_max"],
            MLP_hidden_layers_widths=w["mlp"], HPD_hidden_layers_widths=w["hpd"], HPD_out_features=w["T"],
            feature_dim=w["F"], topk_k=w["K"], should_keep_topk_only=w["topk_only"])
        h, wd = w["lattice_hw"]
        m = max(h, wd) - 1
        net.set_coord_bounds((0.0, 0.0), ((h - 1) / m, (wd - 1) / m))
        self.opt = FusedAdam(                                                       # functions.py:96-127, one launch
            [{"params": net.encoding.parameters(), "lr": w["lr"]["encoding"], "weight_decay": w["wd"]["encoding"]},
             {"params": net.HPD.parameters(), "lr": w["lr"]["hpd"], "weight_decay": w["wd"]["hpd"]},
             {"params": net.mlp.parameters(), "lr": w["lr"]["mlp"], "weight_decay": w["wd"]["mlp"]}],
            betas=(0.9, 0.99), eps=1e-15)
        dp.enable_gradient_allreduce()      # N > 1: column-sum + flat-gradient exchanges, node-parallel HPD
        self.strong = w["scaling"] == "strong"
        if self.strong:
            # the step's points are SPLIT over the ranks: every rank draws the same global batch and keeps its shard
            x_np, y_np = make_inputs(w, 65535, 0)
            a, b = dp.shard_bounds(w["P"], rank, world)
            x_np, y_np = np.ascontiguousarray(x_np[a:b]), np.ascontiguousarray(y_np[a:b])
            self.total_points = w["P"]
        else:
            x_np, y_np = make_inputs(w, 65535, rank)
            self.total_points = w["P"] * world
        self.local_points = x_np.shape[0]
        self.x_host, self.y_host = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(y_np).pin_memory()
        self.x_dev, self.y_dev = self.x_host.to(dev), self.y_host.to(dev)
        self.loss_host = torch.zeros(1).pin_memory()
        self.rows = 4 * self.total_points
        self._loss_fn = fused_total_loss

    def step(self, x, y):
        w = self.w
        self.opt.zero_grad(set_to_none=True)
        rgb, probs, idx, _ = self.net(x, 1.0)
        colsum = self.dp.all_reduce_colsum(probs.colsum) if self.world > 1 else probs.colsum
        loss, _, _ = self._loss_fn(rgb, y, colsum, self.rows, w["gamma"], w["epsilon"], w["l_mse"], w["l_js_kl"])
        loss.backward()
        self.opt.step()
        return loss

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_step(self):
        """The module API end to end: pinned host batch in, loss out, every step."""
        torch = self.torch
        loss = self.step(self.x_host.to(self.dev, non_blocking=True), self.y_host.to(self.dev, non_blocking=True))
        self.loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()


def verify_peer_allreduce(torch, dist, dp, dev, sizes):
    """One peer-kernel vs NCCL comparison per buffer size in use, bit for bit (integer-valued floats: every summation
    order gives the same bits), and identical results on every rank."""
    comm = dp.peer_allreduce_for()
    if comm is None:
        return {"verified": False, "why": "peer all-reduce unavailable (NCCL all_reduce in use)"}
    world, rank = dist.get_world_size(), dist.get_rank()
    ok, checked = True, []
    for n in sorted({int(s) for s in sizes if 0 < int(s) * 4 <= dp.PEER_MAX_BYTES}):
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        t = torch.randint(-1000, 1000, (n,), generator=g, device=dev).float()
        a = dp.all_reduce_sum(t)
        b = t.clone()
        dist.all_reduce(b)
        same = bool(torch.equal(a, b))
        digest = torch.stack([a.double().sum(), (a.double() * torch.arange(n, device=dev)).sum()])
        all_d = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(all_d, digest)
        same = same and all(bool(torch.equal(all_d[0], d)) for d in all_d)
        ok = ok and same
        checked.append(n)
    comm.check()
    return {"verified": bool(ok) and bool(checked), "floats": checked}


def measure(torch, dist, R, steps, warmup, *, profile, eager_e2e, graph, flush):
    """Device-timed steps (+ optional per-kernel profile and end-to-end passes) of Runner R; returns a dict."""
    from collision_handling_in_instantngp_b200 import _lib, launch_count
    from collision_handling_in_instantngp_b200 import ops as _ops
    from collision_handling_in_instantngp_b200.trainer import GraphedTrainer
    w, dev = R.w, R.dev
    out = {}
    for _ in range(max(warmup, MIN_WARMUP)):
        R.step(R.x_dev, R.y_dev)
    R.barrier()
    n0 = launch_count()
    R.step(R.x_dev, R.y_dev)
    torch.cuda.synchronize()
    out["launches_per_step"] = launch_count() - n0

    if eager_e2e:      # what the reference's own train_step gets from the drop-in module: eager, host inputs
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            R.host_step()
        R.barrier()
        out["e2e_eager_ms"] = (time.perf_counter() - t0) * 1e3 / steps

    if profile:        # per-kernel device times (CUDA events around every C-ABI call, serial schedule) -> dominant kernel
        prof = CallProfiler(torch)
        _ops.stream_bwd_stats(reset=True)
        _lib.PROFILER = prof
        _ops.CONCURRENT = False
        for _ in range(profile):
            flush.zero_()
            R.step(R.x_dev, R.y_dev)
        _ops.CONCURRENT = True
        _lib.PROFILER = None
        out["profile"] = prof.summary()
        out["profile_steps"] = profile
        # tiles of the streaming backward's two dense passes and how many of them issued their second product (it is
        # skipped for all-zero E tiles: the executed tensor work depends on the data)
        out["stream_bwd_stats"] = _ops.stream_bwd_stats(reset=True)
    st = R.net.last_state
    out["lat"] = st.lat
    ids = st.node_ids_all if st.shard is not None else st.node_ids
    out["listed_nodes"] = (st.shard[1] if st.shard is not None else (None if ids is None else int(ids.shape[0])))
    out["hpd_rows_this_rank"] = None if st.node_ids is None else int(st.node_ids.shape[0])
    R.barrier()

    trainer = None
    if graph:          # the public fast path for small lattices: the whole step captured once in a CUDA graph
        trainer = GraphedTrainer(R.net, R.opt, points=R.local_points, gamma=w["gamma"], epsilon=w["epsilon"],
                                 l_mse=w["l_mse"], l_js_kl=w["l_js_kl"], warmup_steps=3, sample_x=R.x_dev,
                                 sample_y=R.y_dev)
        for _ in range(max(warmup, 3)):
            trainer.replay()
        torch.cuda.synchronize()
    out["launch_mode"] = "cuda_graph" if trainer is not None else "eager"

    # ---- the device-timed region: inputs resident in HBM, CUDA events around every step, L2 flushed between ----
    R.barrier()
    evs = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if trainer is not None:
            trainer.replay()
        else:
            R.step(R.x_dev, R.y_dev)
        b.record()
        evs.append((a, b))
    R.barrier()
    out["dev_ms"] = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # ---- end to end through the public API with HOST buffers (wall clock) ----
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        if trainer is not None:
            trainer.step_pipelined(R.x_host, R.y_host)
        else:
            R.host_step()
    if trainer is not None:
        trainer.flush()
    R.barrier()
    out["e2e_ms"] = (time.perf_counter() - t0) * 1e3 / steps
    out["e2e_path"] = (
        "trainer.GraphedTrainer.step_pipelined(x_host, y_host): pinned host batch copied into the idle one of two static "
        "buffer sets on a copy stream, CUDA-graph replay of forward+loss+backward+Adam, loss copied to pinned host memory "
        "and returned one step later, every step; flush() inside the timed region" if trainer is not None else
        "module API, eager: net(x) / loss / backward / FusedAdam with pinned host inputs copied in and the loss read back "
        "every step")
    out["h2d_bytes"] = int(R.x_host.numel() * 4 + R.y_host.numel() * 4)
    if trainer is not None:
        trainer.check_errors()
        del trainer
    R.net.check_errors()
    return out


def roofline_of(m, w, steps_profiled, workload):
    per_name = {}
    for (name, key), (t, n) in m["profile"].items():
        bound, amount = algorithmic_cost(name, key, w, m["lat"])
        e = per_name.setdefault(name, dict(ms=0.0, n=0, amount=0.0, bound=bound))
        e["ms"] += t; e["n"] += n; e["amount"] += amount * n
    total_kernel_ms = sum(e["ms"] for e in per_name.values())
    tname, te = max(per_name.items(), key=lambda kv: kv[1]["ms"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # a kernel timed inside a long step sees the sustained clock; short kernels the burst figure
    long_kernel = te["ms"] / te["n"] > 50.0
    tc_peak = peaks.get("bf16_tflops_sustained" if long_kernel else "bf16_tflops", 1590.0)
    per_launch_s = te["ms"] / te["n"] / 1e3
    per_launch_amount = te["amount"] / te["n"]
    if te["bound"] == "hbm":
        achieved, peak, unit = per_launch_amount / per_launch_s / 1e9, hbm_peak, "GB/s"
    elif te["bound"] == "nvlink":
        achieved, peak, unit = per_launch_amount / per_launch_s / 1e9, 900.0, "GB/s"
    else:
        achieved, peak, unit = per_launch_amount / per_launch_s / 1e12, tc_peak, "TFLOP/s"
    return {"bound": te["bound"], "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
            "traffic": NCU_TRAFFIC.get((workload, tname)), "kernel": tname,
            "useful_tflops": useful_tflops(tname, achieved, w), "launches_per_step": te["n"] / steps_profiled,
            "share_of_step_kernel_time": te["ms"] / total_kernel_ms,
            "peak_source": ("measured" if peaks else "fallback") + (
                " (sustained bf16: the kernel runs > 50 ms per launch)" if te["bound"] == "tensor" and long_kernel else
                " (burst bf16)" if te["bound"] == "tensor" else ""),
            "kernels_ms_per_step": {k: round(v["ms"] / steps_profiled, 5) for k, v in
                                    sorted(per_name.items(), key=lambda kv: -kv[1]["ms"])}}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from collision_handling_in_instantngp_b200 import dp

    name = args.workload
    w = dict(WORKLOADS[name])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)               # > 126 MB L2
    R = Runner(torch, dist, name, w, rank, world, dev)
    graph = (not args.eager) and w["T"] <= 4096
    peer = None
    if world > 1:
        n_params = sum((p.numel() + 3) & ~3 for p in R.net._parameters_flat())
        peer = verify_peer_allreduce(torch, dist, dp, dev, [w["L"] * (w["K"] if w["topk_only"] else w["T"]), n_params])

    sampler = ClockSampler(local)
    sampler.start()
    prof_steps = min(args.steps, 3 if w["P"] * w["T"] >= 2 ** 34 else args.steps)
    m = measure(torch, dist, R, args.steps, args.warmup, profile=prof_steps, eager_e2e=graph, graph=graph, flush=flush)
    clocks = sampler.stop()
    wp = dict(w, P=R.local_points, _active_nodes=m["hpd_rows_this_rank"])
    st = m.get("stream_bwd_stats")
    if st and (st[0] + st[3]) > 0:
        # products per tile, averaged over both passes: 1 (screening) + 2 (rest of the logits) + 3 (second product)
        wp["_stream_products_per_tile"] = (st[0] + st[3] + 2.0 * (st[1] + st[4]) + 3.0 * (st[2] + st[5])) / (st[0] + st[3])
    roofline = roofline_of(m, wp, prof_steps, name)
    if st and (st[0] + st[3]) > 0:
        if roofline["kernel"] in ("gngf_hpd_stream_bwd", "gngf_hpd_stream_bwd_nodes") and roofline["unit"] == "TFLOP/s":
            # what a pass that skips nothing (6 split products per tile and pass) would have to sustain to take the same time
            roofline["dense_equivalent_tflops"] = roofline["achieved"] * 6.0 / wp["_stream_products_per_tile"]
        roofline["stream_bwd_tiles"] = {"dh_pass": st[0], "dh_all_logit_products": st[1], "dh_second_product_issued": st[2],
                                        "dw3_pass": st[3], "dw3_all_logit_products": st[4],
                                        "dw3_second_product_issued": st[5], "profiled_steps": prof_steps,
                                        "products_per_tile": wp["_stream_products_per_tile"]}
    dev_ms, e2e_ms, e2e_eager_ms = m["dev_ms"], m["e2e_ms"], m.get("e2e_eager_ms", m["e2e_ms"])
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, e2e_eager_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, e2e_eager_ms = float(t[0]), float(t[1]), float(t[2])
        dp.check_exchanges()
    total = R.total_points
    lat = m["lat"]
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": total / (dev_ms / 1e3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, MIN_WARMUP), "ms_per_step": dev_ms, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(name, w, world),
            "run": {"points_this_rank": R.local_points, "launch_mode": m["launch_mode"],
                    "l2": "flushed between timed steps (256 MiB write)", "lattice_nodes": lat.num_nodes,
                    "level_nodes": lat.num_level_nodes, "hpd_nodes_listed": m["listed_nodes"] or lat.num_nodes,
                    "hpd_rows_rank0": m["hpd_rows_this_rank"] or lat.num_nodes,
                    "node_parallel_hpd": bool(world > 1 and R.net.last_state.shard is not None)},
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": 4, "path": m["e2e_path"],
                    "eager_module_api": {"value": total / (e2e_eager_ms / 1e3), "unit": "samples/s",
                                         "ms_per_step": e2e_eager_ms,
                                         "path": "net(x) / loss / backward / Adam driven eagerly from Python, as the "
                                                 "reference's train_step drives the drop-in module"}},
            "gpu_launches": int(m["launches_per_step"] * args.steps), "gpu_launches_per_step": int(m["launches_per_step"]),
            "clocks": clocks, "roofline": roofline,
        }
        if peer is not None:
            line["peer_allreduce_verified"] = peer["verified"]
            line["peer_allreduce_check"] = peer

    # ---- N = 1: the other named configurations in the same process, then the reference legs ----
    if world == 1 and not args.no_secondary:
        del R, m
        gc.collect()
        torch.cuda.empty_cache()
        secondary = {}
        for sname in [s for s in args.secondary.split(",") if s and s != name]:
            sw = dict(WORKLOADS[sname])
            S = Runner(torch, dist, sname, sw, 0, 1, dev)
            sgraph = sw["T"] <= 4096
            ssteps = args.steps if sgraph else min(args.steps, 5)
            sm = measure(torch, dist, S, ssteps, 3, profile=0, eager_e2e=False, graph=sgraph, flush=flush)
            secondary[sname] = {"names": sw["names"], "value": S.total_points / (sm["dev_ms"] / 1e3), "unit": "samples/s",
                                "ms_per_step": sm["dev_ms"], "steps": ssteps, "launch_mode": sm["launch_mode"],
                                "e2e": {"value": S.total_points / (sm["e2e_ms"] / 1e3), "unit": "samples/s",
                                        "ms_per_step": sm["e2e_ms"], "h2d_bytes_per_step": sm["h2d_bytes"],
                                        "d2h_bytes_per_step": 4},
                                "config": static_config(sname, sw, 1), "lattice_nodes": sm["lat"].num_nodes}
            del S, sm
            gc.collect()
            torch.cuda.empty_cache()
        line["secondary"] = secondary
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        gc.collect()
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_reference_record(w, steps=3, warmup=1, sample_P=args.cpu_sample)
        if reference_available():
            pts = gpu_reference_points(w)
            r = time_reference(w, "cuda", pts, steps=5, warmup=2)
            if r and "error" not in r:
                line["gpu_eager_reference"] = {
                    "value": r["samples_per_s"], "unit": "samples/s", "ms_per_step": r["sec_per_step"] * 1e3,
                    "sample": f"{pts} of the workload's {w['P']} coordinates per step (largest power of two whose "
                              f"(rows, T) tensors fit: peak {r.get('peak_mem_gb', 0):.1f} GB), 5 steps after 2 warm-up",
                    "what": "the UNMODIFIED reference (baseline/_ref) on cuda:0: its ATen eager path, same step",
                    "speedup_of_value": line["value"] / r["samples_per_s"]}
            else:
                line["gpu_eager_reference"] = {"unavailable": (r or {}).get("error", "no output")}
    if rank == 0:
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        teardown(torch, dist)


def teardown(torch, dist):
    """Leave the process group cleanly: graphs and symmetric-memory users are gone by now (measure() deletes its
    trainer), so the communicator can be destroyed.  A watchdog exits the process if the teardown does not return (a
    captured NCCL collective was once observed to hang it)."""
    sys.stdout.flush()
    sys.stderr.flush()
    gc.collect()
    torch.cuda.synchronize()
    timer = threading.Timer(30.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        timer.cancel()


JSON_OUT = sys.stdout
MIN_WARMUP = 3


def _keep_stdout_for_the_json_line():
    """Libraries write to file descriptor 1 (NCCL prints its version banner there): route fd 1 to stderr and keep the
    original stdout for the one JSON line."""
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _keep_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default="cfg2,cfg3", help="other configurations measured in the same run at N = 1")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="points per CPU-baseline step (0: the workload's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not capture the timed step in a CUDA graph")
    ap.add_argument("--min-warmup", type=int, default=3, help="lower bound on the warm-up steps (builder runs of the "
                    "50 s / step T = 2^19 configuration use 1)")
    args = ap.parse_args()
    global MIN_WARMUP
    MIN_WARMUP = max(1, args.min_warmup)
    w = dict(WORKLOADS[args.workload])

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        # the reference's CPU path on the box's host cores, on a bounded sample of OUR arm's configuration (the sample
        # is sized so that W + K steps end within a few minutes: cpu_sample in WORKLOADS)
        steps, warmup = args.steps, args.warmup
        rec = cpu_reference_record(w, steps=steps, warmup=warmup, sample_P=args.cpu_sample)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
            "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(args.workload, w, args.gpus),
            "cpu_baseline": rec,
            "e2e": {"value": rec["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }), file=JSON_OUT, flush=True)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
