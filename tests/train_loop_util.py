"""Test-side restatement of the reference's training driver for one parameter set: the batch loop of
`train_step` (functions.py:139-355), the optimizer of `get_optimizer` (functions.py:96-127) and the per-epoch
PSNR of `grid_search_loop` (functions.py:653-692, 130-136).  It drives the drop-in module exactly as the
reference's loop does (same batch slicing of the shuffled pixel order, same loss assembly, same Adam groups), so
that PSNR trajectories can be compared with `oracle/run_reference_training.py`'s golden run on the GPU box,
where the reference tree itself is not available."""
import numpy as np
import torch

from parity_util import RefLoss

# params.py:26-51 + grid-search ID 4061 (README.md:15-18)
ID_4061 = dict(T=256, L=4, n_min=8, n_max=32, F=2, K=4, hpd=[32, 64, 128], mlp=[64, 64], batch_size=1 / 3,
               gamma=-2.0, epsilon=1.0, l_mse=1.0, l_js_kl=1.0, l_collisions=1e-3, topk_only=False,
               encoding_lr=1e-4, HPD_lr=1e-3, MLP_lr=1e-3, encoding_wd=0.0, HPD_wd=1e-6, MLP_wd=1e-6)


def image_dataset(img_u8, device):
    """main.py:42-51 / utils.py:32-66: X = (row, col) / (max(w,h)-1), Y = rgb / 255."""
    h, w = img_u8.shape[:2]
    X = np.stack(np.meshgrid(range(h), range(w), indexing="ij"), -1).reshape(-1, 2)
    x = torch.tensor(X).float() / (max(w, h) - 1)
    y = torch.tensor(img_u8.reshape(-1, 3) / 255).float()
    return x.to(device), y.to(device), h, w


def calc_psnr(pred, target):
    mse = np.square(pred - target).mean()                      # functions.py:134-136
    return 20 * np.log10(np.max(target)) - 10 * np.log10(mse)


def make_optimizer(net, c):
    return torch.optim.Adam(
        [{"params": net.encoding.parameters(), "lr": c["encoding_lr"], "weight_decay": c["encoding_wd"]},
         {"params": net.HPD.parameters(), "lr": c["HPD_lr"], "weight_decay": c["HPD_wd"]},
         {"params": net.mlp.parameters(), "lr": c["MLP_lr"], "weight_decay": c["MLP_wd"]}],
        betas=(0.9, 0.99), eps=1e-15)


def train_epochs(net, x, y, img_u8, shuffled, reordered, epochs, c=ID_4061):
    dev = x.device
    h, w = img_u8.shape[:2]
    shape = w * h
    opt = make_optimizer(net, c)
    loss_fn = RefLoss(c["gamma"], c["epsilon"])
    pct = c["batch_size"]
    num_batches = int(np.ceil(shape / (shape * pct)))
    prev_coll, prev_min = torch.tensor([], device=dev), torch.tensor([], device=dev)
    shuffled = torch.as_tensor(shuffled, device=dev).long()
    reordered = torch.as_tensor(reordered, device=dev).long()
    K = c["K"]
    hist = {"psnr": [], "loss": [], "mse": [], "collisions": []}
    for _ in range(epochs):
        net.train()
        outputs = torch.empty((shape, 3), device=dev)
        indices = torch.empty((shape, c["L"], 4, int(K * 1 / pct)), device=dev)   # functions.py:179: torch.empty, as is
        losses, mses = [], []
        for b in range(num_batches):
            start, stop = b * int(pct * shape), (b + 1) * int(pct * shape)
            sel = shuffled[start:stop]
            bx, by = x[sel], y[sel]
            opt.zero_grad()
            rgb, probs, idx, _ = net(bx, pct)
            outputs[start:stop] = rgb.detach()
            indices[start:stop, ..., b * K:(b + 1) * K] = idx
            mse, kl, coll_l = loss_fn(rgb, by, probs.shape[-1], probs, prev_coll, prev_min)
            loss = c["l_mse"] * mse
            loss = loss + ((c["l_js_kl"] * kl) + (c["l_collisions"] * coll_l if coll_l.nelement() != 0 else 1)).sum(0)
            losses.append(float(loss.detach()))
            mses.append(float(mse.detach()))
            loss.backward()
            opt.step()
        output = outputs[reordered]
        prev_coll, prev_min = net.calc_hash_collisions(indices[reordered])
        img = (output * 255).reshape(h, w, 3).int().cpu().numpy()
        hist["psnr"].append(calc_psnr(img, img_u8))
        hist["loss"].append(float(np.mean(losses)))
        hist["mse"].append(float(np.mean(mses)))
        hist["collisions"].append(prev_coll.cpu().numpy())
    return hist
