"""Helpers shared by the GPU parity tests, smoke() and bench.py: load golden parameters into the drop-in
module and run one training step (forward, the reference's loss assembly, backward)."""
import numpy as np
import torch


def build_net(g, **over):
    """Constructs the drop-in GeneralNeuralGaugeFields for a golden case's configuration."""
    from collision_handling_in_instantngp_b200 import models as M
    c = dict(g["cfg"])
    c.update(over)
    M.DEFAULT_FLAGS.should_use_hash_function = c["use_hash"]
    M.DEFAULT_FLAGS.should_softmax_topk_features = c["mix_mode"]
    M.DEFAULT_FLAGS.should_leaky_relu = c["leaky"]
    M.DEFAULT_FLAGS.should_inplace_scatter = c.get("inplace", True)
    net = M.GeneralNeuralGaugeFields(
        input_dim=2, hash_table_size=c["T"], num_levels=c["L"], n_min=c["n_min"], n_max=c["n_max"],
        MLP_hidden_layers_widths=c["mlp"], HPD_hidden_layers_widths=c["hpd"], HPD_out_features=c["T"],
        feature_dim=c["F"], topk_k=c["K"], should_keep_topk_only=c["topk_only"], should_bw=c["bw"])
    load_golden_into(net, g)
    return net


def reset_flags():
    from collision_handling_in_instantngp_b200 import models as M
    M.DEFAULT_FLAGS.should_use_hash_function = False
    M.DEFAULT_FLAGS.should_softmax_topk_features = True
    M.DEFAULT_FLAGS.should_leaky_relu = False
    M.DEFAULT_FLAGS.should_inplace_scatter = True


def load_golden_into(net, g):
    sd = net.state_dict()
    new = {}
    for k, v in sd.items():
        key = "param." + k
        new[k] = torch.from_numpy(g[key]).to(v.device) if key in g else v
    net.load_state_dict(new)


class RefLoss(torch.nn.Module):
    """Test-side restatement of utils.py:78-174 (same torch modules and op order as the reference), used to
    drive the drop-in module exactly the way functions.py:227-245 does."""

    def __init__(self, gamma, epsilon, delta=1):
        super().__init__()
        self.mse = torch.nn.MSELoss()
        self.kl = torch.nn.KLDivLoss(reduction="batchmean")
        self.gamma, self.epsilon, self.delta = gamma, epsilon, delta

    def forward(self, pred, labels, N, prob, collisions, min_possible):
        mse = self.mse(pred, labels)
        coll = collisions / (min_possible + self.delta)
        levels = torch.stack([self.level(N, prob[:, l, :], prob.shape[0] * prob.shape[2]) for l in range(prob.shape[1])])
        return mse, levels, coll

    def level(self, N, p, div):
        dev = None
        q = torch.ones(N, device=dev) / float(N)
        p_out = p.sum(0).sum(0) / div
        q = q.to(p_out.device)
        kl = self.kl(p_out.log(), q)
        m = (p_out + q) / 2
        js = (self.kl(p_out.log(), m) + self.kl(q.log(), m)) / 2
        return -(self.gamma + self.epsilon) * js + self.epsilon * kl


def run_step(net, g, materialize=False):
    """One step on the golden inputs: returns numpy outputs + parameter gradients."""
    c = g["cfg"]
    dev = next(net.parameters()).device
    x = torch.from_numpy(g["x"]).to(dev)
    y = torch.from_numpy(g["y"]).to(dev)
    net.zero_grad()
    rgb, probs, idx, _ = net(x, 1.0)
    out = {"rgb": rgb.detach().cpu().numpy(), "idx": idx.cpu().numpy()}
    if c["use_hash"]:
        loss = c["l_mse"] * torch.nn.functional.mse_loss(rgb, y)
    else:
        loss_fn = RefLoss(c["gamma"], c["epsilon"])
        if "collisions" in g:
            coll = torch.from_numpy(g["collisions"]).to(dev)
            minp = torch.from_numpy(g["min_possible"]).to(dev)
        else:
            coll, minp = torch.tensor([], device=dev), torch.tensor([], device=dev)
        pr = probs.materialize() if materialize else probs
        mse, levels, coll_l = loss_fn(rgb, y, probs.shape[-1], pr, coll, minp)
        loss = c["l_mse"] * mse
        loss = loss + ((c["l_js_kl"] * levels) + (c["l_collisions"] * coll_l if coll_l.nelement() != 0 else 1)).sum(0)
        out.update(mse=float(mse), kl_levels=levels.detach().cpu().numpy(),
                   pbar=(probs.colsum / (probs.shape[0] * probs.shape[2])).detach().cpu().numpy())
    loss.backward()
    out["loss"] = float(loss)
    out["grads"] = {k: v.grad.detach().cpu().numpy() for k, v in net.named_parameters() if v.grad is not None}
    st = net.last_state
    out["state"] = st
    return out


def decoder_masks(state):
    """The decoder's ReLU pattern as the tensor-core forward recorded it: [None, (P,64) bool, (P,64) bool]
    (k6_mlp_tc.cu: words [layer*2 + half], bit j = unit half*32 + j), or None when the fp32 decoder ran."""
    if getattr(state, "mlp_masks", None) is None or not state.mlp_tc:
        return None
    m = state.mlp_masks.cpu().numpy().view(np.uint32)                 # (P,4)
    bits = ((m[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)   # (P,4,32)
    return [None, bits[:, 0:2].reshape(-1, 64), bits[:, 2:4].reshape(-1, 64)]
