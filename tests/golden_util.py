"""Loads tests/golden/*.npz (written by oracle/make_goldens.py from the unmodified reference)."""
import ast
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GNGF_CASES = ["cfg2_small", "cfg2_topk_only", "cfg2_epoch1", "mix_weighted_avg", "mix_raw", "k1", "k20",
              "bw_leaky", "l16_t1024", "l8_t4096_topk_only", "js_only", "kl_only", "scatter_none",
              "scatter_none_topk_only", "counts", "counts_l16"]
ALL_CASES = GNGF_CASES + ["hash_mode", "counts_hash"]
COUNTS_CASES = ["counts", "counts_hash", "counts_l16"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["cfg"] = ast.literal_eval(str(g["cfg"]))
    return g


def params_of(g, dtype=np.float32):
    """golden -> the oracle's params dict."""
    c = g["cfg"]
    p = {"tables": [g[f"param.encoding._hash_tables.{l}.weight"].astype(dtype) for l in range(c["L"])]}
    n_mlp = len(c["mlp"]) + 1
    p["mlp_w"] = [g[f"param.mlp.{i}.0.weight"].astype(dtype) for i in range(n_mlp)]
    p["mlp_b"] = [g[f"param.mlp.{i}.0.bias"].astype(dtype) for i in range(n_mlp)]
    if not c["use_hash"]:
        n_hpd = len(c["hpd"]) + 1
        p["hpd_w"] = [g[f"param.HPD.module_list.{i}.0.weight"].astype(dtype) for i in range(n_hpd)]
        p["hpd_b"] = [g[f"param.HPD.module_list.{i}.0.bias"].astype(dtype) for i in range(n_hpd)]
    return p


def oracle_cfg(g):
    c = g["cfg"]
    return {"n_ls": g["n_ls"].astype(np.int32), "table_size": c["T"], "topk_k": c["K"], "mix_mode": c["mix_mode"],
            "use_hash": c["use_hash"], "leaky": c["leaky"], "topk_only": c["topk_only"],
            "drop_topk_adjoint": c.get("inplace", True) is None}


def golden_counts(g):
    """The reference's counts_per_level (list per level of {slot: count}) as stored by make_goldens.py."""
    return [dict(zip(g[f"counts_keys_{l}"].tolist(), g[f"counts_vals_{l}"].tolist())) for l in range(g["cfg"]["L"])]


def loss_cfg(g):
    c = g["cfg"]
    return {"gamma": c["gamma"], "epsilon": c["epsilon"], "l_mse": c["l_mse"], "l_js_kl": c["l_js_kl"],
            "l_collisions": c["l_collisions"]}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))
