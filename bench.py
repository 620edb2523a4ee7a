#!/usr/bin/env python
"""bench.py -- train samples/s of the GNGF hot path (forward + loss + backward + Adam) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

One "step" = one pass of the hot path over one batch of synthetic coordinates:
GeneralNeuralGaugeFields.forward (models.py:394-484) -> the reference's loss assembly (utils.py:78-174,
functions.py:243-245) -> backward -> Adam (functions.py:96-127).  The default workload is BASELINE.json
configs[1] ("cfg2": the shape of strawberry.jpeg with grid-search ID 4061: P = 57 404 pixels per batch, L = 4
levels n = [8,12,20,32], T = 256 slots, K = 4, F = 2, HPD 2-32-64-128-256, decoder 8-64-64-3) on synthetic data
(pixel-lattice coordinates in a seeded random order, uniform random targets, random-init weights).

JSON line (rank 0): see README/DESIGN.md.  `value` is device time with inputs resident in HBM (CUDA events per
step, L2 flushed between steps); `e2e` is the same step driven through the module API with HOST (pinned)
inputs copied in and the loss read back every step; `roofline` describes the dominant kernel of the step;
`cpu_baseline` is the oracle port (numpy restatement of the reference) timed on this box's host cores.
`--impl reference` times that CPU port alone (the reference is PyTorch-on-CPU code that cannot travel to the
GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: strawberry.jpeg + param ID 4061 (README.md:15-18, params.py:26-51)
    "cfg2": dict(P=57404, L=4, n_min=8, n_max=32, T=256, K=4, F=2, hpd=[32, 64, 128], mlp=[64, 64],
                 lattice_hw=(508, 339), topk_only=False, gamma=-2.0, epsilon=1.0, l_mse=1.0, l_js_kl=1.0,
                 lr=dict(encoding=1e-4, hpd=1e-3, mlp=1e-3), wd=dict(encoding=0.0, hpd=1e-6, mlp=1e-6),
                 cpu_sample=16384),
    # BASELINE.json configs[2]: macaw.jpg (508x339 = 172 212 px < 2^18: the whole image is one batch), 16 levels,
    # table size 2^19, top-k-only probabilities (the full distribution would be 134 MB per sample)
    "cfg3": dict(P=172212, L=16, n_min=16, n_max=508, T=2 ** 19, K=4, F=2, hpd=[32, 64, 128], mlp=[64, 64],
                 lattice_hw=(508, 339), topk_only=True, gamma=-2.0, epsilon=1.0, l_mse=1.0, l_js_kl=1.0,
                 lr=dict(encoding=1e-4, hpd=1e-3, mlp=1e-3), wd=dict(encoding=0.0, hpd=1e-6, mlp=1e-6),
                 cpu_sample=4),
    # same image shape with a mid-size table (fits the dense path as well; used to compare the two HPD paths)
    "cfg3_t14": dict(P=172212, L=16, n_min=16, n_max=508, T=2 ** 14, K=4, F=2, hpd=[32, 64, 128], mlp=[64, 64],
                     lattice_hw=(508, 339), topk_only=True, gamma=-2.0, epsilon=1.0, l_mse=1.0, l_js_kl=1.0,
                     lr=dict(encoding=1e-4, hpd=1e-3, mlp=1e-3), wd=dict(encoding=0.0, hpd=1e-6, mlp=1e-6),
                     cpu_sample=64),
    # BASELINE.json configs[3]: synthetic 8192 x 8192 lattice, 2^22 points per step, 16 levels x 2 features
    "cfg4_t14": dict(P=2 ** 22, L=16, n_min=16, n_max=8192, T=2 ** 14, K=4, F=2, hpd=[32, 64, 128], mlp=[64, 64],
                     lattice_hw=(8192, 8192), topk_only=True, gamma=-2.0, epsilon=1.0, l_mse=1.0, l_js_kl=1.0,
                     lr=dict(encoding=1e-4, hpd=1e-3, mlp=1e-3), wd=dict(encoding=0.0, hpd=1e-6, mlp=1e-6),
                     cpu_sample=64),
}


def make_inputs(w, seed, rank=0):
    """Synthetic batch of the workload's shape: coordinates are pixel-lattice points (row, col)/(max(h,w)-1) as in
    main.py:42-51, drawn without replacement in a seeded random order; targets uniform in [0,1)."""
    h, wd = w["lattice_hw"]
    rng = np.random.default_rng(seed + 7919 * rank)
    flat = rng.permutation(h * wd)[: w["P"]]
    x = np.stack([flat // wd, flat % wd], 1).astype(np.float32) / np.float32(max(h, wd) - 1)
    y = rng.random((w["P"], 3), dtype=np.float32)
    return x, y


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's path restated on PyTorch-CPU (oracle/torch_port.py), forward + loss + backward + Adam
# ------------------------------------------------------------------------------------------------------------
def time_cpu_port(w, sample_P, seed, steps, warmup):
    """Seconds per step of the PyTorch-CPU port of the reference (oracle/torch_port.py: the same ATen operations
    and autograd the reference issues, all host threads) on `sample_P` coordinates of the workload."""
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


def cpu_threads():
    import torch
    return int(torch.get_num_threads())


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
    return ()


def algorithmic_cost(name, key, w, lat):
    """(bound, amount per launch, unit): ALGORITHMIC bytes (hbm) or FLOPs (tensor) of one launch."""
    P, L, F, K, T = w["P"], w["L"], w["F"], w["K"], w["T"]
    U, S = lat.num_nodes, lat.num_level_nodes
    Ua = w.get("_active_nodes") or U      # rows of the HPD chain (ops.active_nodes: the nodes the batch touches)
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
        # two fused passes (dh, dW3), each: logits recomputed (3 split products) + second product (3 split products);
        # useful work = the two gradient products, 2 * 2*U*T*kd
        return "tensor", 12 * 2.0 * Ua * T * kd
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
        "gngf_split_bf16x3": 0.0,
    }
    return "hbm", float(table.get(name, 0))


def useful_tflops(name, achieved, w):
    """The share of the executed tensor-core FLOPs that the fp32 algorithm asks for (one pass per product)."""
    if name in ("gngf_hpd_stream_fwd", "gngf_tc_gemm_bf16x3"):
        return achieved / 6
    if name == "gngf_hpd_stream_fwd_refined":
        return achieved / 3
    if name in ("gngf_hpd_stream_bwd", "gngf_hpd_stream_bwd_nodes"):
        return achieved / 6                       # 12 executed passes for the 2 gradient products
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


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    from collision_handling_in_instantngp_b200 import _lib, dp, launch_count
    from collision_handling_in_instantngp_b200.loss import fused_total_loss as total_loss
    from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
    from collision_handling_in_instantngp_b200.optim import FusedAdam
    from collision_handling_in_instantngp_b200.trainer import GraphedTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(65535)

    net = GeneralNeuralGaugeFields(input_dim=2, hash_table_size=w["T"], num_levels=w["L"], n_min=w["n_min"],
                                   n_max=w["n_max"], MLP_hidden_layers_widths=w["mlp"],
                                   HPD_hidden_layers_widths=w["hpd"], HPD_out_features=w["T"], feature_dim=w["F"],
                                   topk_k=w["K"], should_keep_topk_only=w["topk_only"])
    h, wd = w["lattice_hw"]
    m = max(h, wd) - 1
    net.set_coord_bounds((0.0, 0.0), ((h - 1) / m, (wd - 1) / m))
    opt = FusedAdam(                                                            # functions.py:96-127, one launch
        [{"params": net.encoding.parameters(), "lr": w["lr"]["encoding"], "weight_decay": w["wd"]["encoding"]},
         {"params": net.HPD.parameters(), "lr": w["lr"]["hpd"], "weight_decay": w["wd"]["hpd"]},
         {"params": net.mlp.parameters(), "lr": w["lr"]["mlp"], "weight_decay": w["wd"]["mlp"]}],
        betas=(0.9, 0.99), eps=1e-15)
    dp.enable_gradient_allreduce()          # N > 1: one in-place NCCL all-reduce of the flat gradient buffer

    x_np, y_np = make_inputs(w, 65535, rank)
    x_host, y_host = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(y_np).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    loss_host = torch.zeros(1).pin_memory()
    rows = 4 * w["P"] * world

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        rgb, probs, idx, _ = net(x, 1.0)
        colsum = dp.all_reduce_colsum(probs.colsum) if world > 1 else probs.colsum
        loss, _, _ = total_loss(rgb, y, colsum, rows, w["gamma"], w["epsilon"], w["l_mse"], w["l_js_kl"])
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)               # > 126 MB L2

    # ---- warm-up (eager) ----
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    barrier()
    n0 = launch_count()
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    launches_per_step = launch_count() - n0

    sampler = ClockSampler(local)
    sampler.start()

    # ---- eager module API end to end (what the reference's own train_step does with the drop-in module) ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        loss = step(xd, yd)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    barrier()
    e2e_eager_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    del loss, xd, yd          # (the last eager loss would keep default-stream AccumulateGrad nodes alive at capture)

    # ---- per-kernel device times (eager pass, CUDA events around every C-ABI call) -> dominant kernel ----
    prof = CallProfiler(torch)
    _lib.PROFILER = prof
    from collision_handling_in_instantngp_b200 import ops as _ops
    _ops.CONCURRENT = False          # serial schedule for this pass only: an event pair around a call on a forked side
    for _ in range(args.steps):      # stream would also time the wait for the fork point
        flush.zero_()
        step(x_dev, y_dev)
    _ops.CONCURRENT = True
    _lib.PROFILER = None
    agg = prof.summary()
    lat = net.last_state.lat
    ids = net.last_state.node_ids
    w = dict(w, _active_nodes=None if ids is None else int(ids.shape[0]))
    barrier()

    # ---- the public fast path: the whole step (our kernels, fused Adam, for N > 1 the two NCCL all-reduces)
    #      captured once in a CUDA graph over static input buffers (trainer.GraphedTrainer) ----
    trainer, launch_mode = None, "eager"
    if not args.eager and w["T"] <= 4096:
        trainer = GraphedTrainer(net, opt, points=w["P"], gamma=w["gamma"], epsilon=w["epsilon"], l_mse=w["l_mse"],
                                 l_js_kl=w["l_js_kl"], warmup_steps=3, sample_x=x_dev, sample_y=y_dev)
        launch_mode = "cuda_graph"
        for _ in range(max(args.warmup, 3)):
            trainer.replay()
        torch.cuda.synchronize()

    # ---- the device-timed region: inputs resident in HBM, CUDA events around every step, L2 flushed between ----
    barrier()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if trainer is not None:
            trainer.replay()
        else:
            step(x_dev, y_dev)
        b.record()
        evs.append((a, b))
    barrier()
    dev_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # ---- end to end through the public API with HOST buffers: every step copies the batch in from pinned host
    #      memory and reads the loss back (wall clock, max over ranks).  GraphedTrainer.step_pipelined double-buffers
    #      the inputs: the copy of batch t+1 overlaps the replay of step t, and the loss of step t is handed to the
    #      caller (as a host float) at step t+1; flush() inside the timed region collects the last one ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if trainer is not None:
            trainer.step_pipelined(x_host, y_host)
        else:
            loss = step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True))
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
    if trainer is not None:
        trainer.flush()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()

    per_name = {}
    for (name, key), (t, n) in agg.items():
        bound, amount = algorithmic_cost(name, key, w, lat)
        e = per_name.setdefault(name, dict(ms=0.0, n=0, amount=0.0, bound=bound))
        e["ms"] += t; e["n"] += n; e["amount"] += amount * n
    total_kernel_ms = sum(e["ms"] for e in per_name.values())
    top = max(per_name.items(), key=lambda kv: kv[1]["ms"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tc_peak = peaks.get("bf16_tflops", 1590.0)
    peak_src = "measured" if peaks else "fallback"
    tname, te = top
    per_launch_s = te["ms"] / te["n"] / 1e3
    per_launch_amount = te["amount"] / te["n"]
    if te["bound"] == "hbm":
        achieved, peak, unit = per_launch_amount / per_launch_s / 1e9, hbm_peak, "GB/s"
    else:
        achieved, peak, unit = per_launch_amount / per_launch_s / 1e12, tc_peak, "TFLOP/s"
    roofline = {"bound": te["bound"], "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get((args.workload, tname)), "kernel": tname,
                "useful_tflops": useful_tflops(tname, achieved, w),
                "launches_per_step": te["n"] / args.steps,
                "share_of_step_kernel_time": te["ms"] / total_kernel_ms, "peak_source": peak_src,
                "kernels_ms_per_step": {k: round(v["ms"] / args.steps, 5) for k, v in
                                        sorted(per_name.items(), key=lambda kv: -kv[1]["ms"])}}

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, e2e_eager_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, e2e_eager_ms = float(t[0]), float(t[1]), float(t[2])

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_P = args.cpu_sample or w["cpu_sample"]
        sec = time_cpu_port(w, sample_P, 65535, steps=3, warmup=1)
        cpu_base = {"value": sample_P / sec, "unit": "samples/s", "cores": cpu_threads(), "kind": "port",
                    "sample": f"{sample_P} of the workload's {w['P']} coordinates per step, 3 steps after 1 warm-up, "
                              "oracle/torch_port.py (PyTorch-CPU restatement of the reference: same ATen ops + autograd) "
                              "forward+loss+backward+Adam"}
    if rank == 0:
        total = w["P"] * world
        line = {
            "metric": "train samples/sec (fwd+bwd GNGF hash encode+MLP)", "value": total / (dev_ms / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "points_per_gpu_step": w["P"], "levels": w["L"],
                       "table_size": w["T"], "topk_k": w["K"], "feature_dim": w["F"], "hpd": [2, *w["hpd"], w["T"]],
                       "decoder": [w["L"] * w["F"], *w["mlp"], 3], "step": "forward+loss+backward+Adam",
                       "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"dp{world}",
                       "launch_mode": launch_mode,
                       "lattice_nodes": lat.num_nodes, "level_nodes": lat.num_level_nodes,
                       "active_nodes": w["_active_nodes"] or lat.num_nodes},
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4), "d2h_bytes_per_step": 4,
                    "path": ("trainer.GraphedTrainer.step_pipelined(x_host, y_host): pinned host batch copied into the idle "
                             "one of two static buffer sets on a copy stream, CUDA-graph replay of "
                             "forward+loss+backward+Adam, loss copied to pinned host memory and returned one step later, "
                             "every step; flush() inside the timed region"
                             if trainer is not None else
                             "module API, eager, pinned host inputs copied in and loss read back every step"),
                    "eager_module_api": {"value": total / (e2e_eager_ms / 1e3), "unit": "samples/s",
                                         "ms_per_step": e2e_eager_ms,
                                         "path": "net(x) / loss / backward / Adam driven eagerly from Python, as the "
                                                 "reference's train_step drives the drop-in module"}},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roofline,
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        # tearing the NCCL communicator down after its collectives were captured in a CUDA graph was observed
        # to hang; the line is out, so leave without the teardown
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


JSON_OUT = sys.stdout


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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="points per CPU-baseline step (0: the workload's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not capture the timed step in a CUDA graph")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        sample_P = args.cpu_sample or w["cpu_sample"]
        sec = time_cpu_port(w, sample_P, 65535, steps=args.steps, warmup=args.warmup)
        val = sample_P / sec
        cores = cpu_threads()
        sample = (f"{sample_P} of the workload's {w['P']} coordinates per step; oracle/torch_port.py (PyTorch-CPU "
                  "restatement of the reference: same ATen ops + autograd, all host threads) forward+loss+backward+Adam")
        print(json.dumps({
            "impl": "reference", "metric": "train samples/sec (fwd+bwd GNGF hash encode+MLP)", "value": val,
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "points_per_step_sample": sample_P, "levels": w["L"],
                       "table_size": w["T"], "topk_k": w["K"], "feature_dim": w["F"],
                       "step": "forward+loss+backward+Adam"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }), file=JSON_OUT, flush=True)
        return
    run_ours(args, w)


if __name__ == "__main__":
    main()
