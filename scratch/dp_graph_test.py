import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
def log(*a):
    print(f"[r{os.environ.get('RANK')}]", *a, file=sys.stderr, flush=True)
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import bench
from collision_handling_in_instantngp_b200 import dp
from collision_handling_in_instantngp_b200.loss import fused_total_loss
from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
w = dict(bench.WORKLOADS["cfg2"]); torch.manual_seed(0)
net = GeneralNeuralGaugeFields(2, w["T"], w["L"], w["n_min"], w["n_max"], w["mlp"], w["hpd"], HPD_out_features=w["T"], topk_k=w["K"])
net.set_coord_bounds((0, 0), (1.0, 338 / 507))
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
params = [p for g in opt.param_groups for p in g["params"]]
x_np, y_np = bench.make_inputs(w, 1, rank)
x, y = torch.from_numpy(x_np).to(dev), torch.from_numpy(y_np).to(dev)
# static flat gradient buffer for the all-reduce
reducer = dp.GradientAllReducer(params)
def step():
    opt.zero_grad(set_to_none=True)
    rgb, probs, idx, _ = net(x, 1.0)
    colsum = dp.all_reduce_colsum(probs.colsum)
    loss, _, _ = fused_total_loss(rgb, y, colsum, 4 * w["P"] * world, -2.0, 1.0)
    loss.backward()
    reducer()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize(); log("eager ok")
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize(); log("side-stream warmup ok")
dist.barrier(); log("barrier ok")
g = torch.cuda.CUDAGraph()
opt.zero_grad(set_to_none=True); net.last_state = None
log("capturing")
with torch.cuda.graph(g, stream=side):
    loss = step()
log("captured")
for _ in range(3): g.replay()
torch.cuda.synchronize(); log("replay ok", float(loss))
t0 = time.perf_counter()
for _ in range(50): g.replay()
torch.cuda.synchronize(); log("ms/step graph", (time.perf_counter() - t0) / 50 * 1e3)
dist.destroy_process_group()
