import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from collision_handling_in_instantngp_b200.loss import total_loss
from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
w = dict(bench.WORKLOADS["cfg2"])
dev = torch.device("cuda")
torch.manual_seed(0)
net = GeneralNeuralGaugeFields(2, w["T"], w["L"], w["n_min"], w["n_max"], w["mlp"], w["hpd"], HPD_out_features=w["T"], topk_k=w["K"])
net.set_coord_bounds((0, 0), (1.0, 338 / 507))
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
x_np, y_np = bench.make_inputs(w, 1)
x, y = torch.from_numpy(x_np).to(dev), torch.from_numpy(y_np).to(dev)
def step(do_opt=True):
    opt.zero_grad(set_to_none=True)
    rgb, probs, idx, _ = net(x, 1.0)
    loss, _, _ = total_loss(rgb, y, probs.colsum, 4 * w["P"], -2.0, 1.0)
    loss.backward()
    if do_opt: opt.step()
    return loss
for _ in range(5): step()
torch.cuda.synchronize()
def timeit(fn, n=20, sync_each=False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync_each: torch.cuda.synchronize()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("step wall ms (no sync each):", timeit(step))
print("step wall ms (sync each):", timeit(step, sync_each=True))
print("step no-opt ms:", timeit(lambda: step(False)))
def fwd_only():
    with torch.no_grad():
        net(x, 1.0)
print("fwd only ms:", timeit(fwd_only))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))
