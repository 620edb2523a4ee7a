"""GPU: PSNR parity on strawberry.jpeg, grid-search ID 4061, same initial weights and pixel order as the
reference run recorded by oracle/run_reference_training.py (unmodified reference, CPU, seed 65535).
Bar (BASELINE.json north_star): PSNR within 0.1 dB."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from train_loop_util import ID_4061, image_dataset, train_epochs

pytestmark = pytest.mark.gpu


def test_psnr_trajectory_matches_reference_run():
    path = os.path.join(GOLDEN_DIR, "ref_trajectory_4061.npz")
    z = np.load(path)
    if "shuffled_indices" not in z.files:
        pytest.skip("golden trajectory was recorded without initial weights")
    img = np.load(os.path.join(GOLDEN_DIR, "strawberry_u8.npz"))["img"]
    from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
    c = ID_4061
    net = GeneralNeuralGaugeFields(input_dim=2, hash_table_size=c["T"], num_levels=c["L"], n_min=c["n_min"],
                                   n_max=c["n_max"], MLP_hidden_layers_widths=c["mlp"],
                                   HPD_hidden_layers_widths=c["hpd"], HPD_out_features=c["T"], feature_dim=c["F"],
                                   topk_k=c["K"], should_keep_topk_only=c["topk_only"])
    sd = net.state_dict()
    for k in sd:
        if "init." + k in z.files:
            sd[k] = torch.from_numpy(z["init." + k]).to(sd[k].device)
    net.load_state_dict(sd)
    x, y, h, w = image_dataset(img, torch.device("cuda"))
    ref_psnr = z["psnr"]
    epochs = len(ref_psnr)
    hist = train_epochs(net, x, y, img, z["shuffled_indices"], z["reordered_indices"], epochs)
    got = np.array(hist["psnr"])
    print("reference PSNR:", np.round(ref_psnr, 4))
    print("this repo PSNR:", np.round(got, 4))
    assert np.abs(got - ref_psnr).max() < 0.1, (got, ref_psnr)
    # the MSE part of the loss follows the reference's as well (the collision term is a constant)
    assert np.abs(np.array(hist["mse"]) - z["mse"]).max() < 2e-3
