"""How much of the cfg2 step is launch / dependency overhead of the graph?  Replay time vs. points per step."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
from collision_handling_in_instantngp_b200.optim import FusedAdam
from collision_handling_in_instantngp_b200.trainer import GraphedTrainer
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for P in (128, 4096, 57404):
    torch.manual_seed(0)
    net = GeneralNeuralGaugeFields(input_dim=2, hash_table_size=256, num_levels=4, n_min=8, n_max=32,
                                   MLP_hidden_layers_widths=[64, 64], HPD_hidden_layers_widths=[32, 64, 128],
                                   HPD_out_features=256, feature_dim=2, topk_k=4)
    net.set_coord_bounds((0.0, 0.0), (1.0, 338 / 507))
    opt = FusedAdam([{"params": net.encoding.parameters(), "lr": 1e-4}, {"params": net.HPD.parameters(), "lr": 1e-3},
                     {"params": net.mlp.parameters(), "lr": 1e-3}], betas=(0.9, 0.99), eps=1e-15)
    x = torch.rand((P, 2), device=dev) * torch.tensor([1.0, 338 / 507], device=dev)
    y = torch.rand((P, 3), device=dev)
    tr = GraphedTrainer(net, opt, points=P, gamma=-2.0, epsilon=1.0, sample_x=x, sample_y=y)
    for _ in range(5):
        tr.replay()
    torch.cuda.synchronize()
    for fl in (False, True):
        ts = []
        for _ in range(30):
            if fl:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); tr.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"P={P:6d} flush={fl}: median {np.median(ts):.1f} us  min {np.min(ts):.1f} us")
    # back-to-back replays (no per-replay sync): steady-state throughput
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        tr.replay()
    b.record(); torch.cuda.synchronize()
    print(f"P={P:6d} back-to-back: {a.elapsed_time(b) * 1e3 / 50:.1f} us/step")
