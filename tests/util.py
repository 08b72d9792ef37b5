import functools
import json
import os

import numpy as np
import torch

from oracle import decoder_oracle as O
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

SPECS = {
    "NetworkSpecs": {"dims": [512] * 8, "dropout": list(range(8)), "dropout_prob": 0.2, "norm_layers": list(range(8)),
                     "latent_in": [4], "xyz_in_all": False, "use_tanh": False, "latent_dropout": False,
                     "weight_norm": True},
    "SubnetSpecs": {"subnet_dims": [512] * 5, "subnet_dropout": list(range(5)), "subnet_norm": list(range(5)),
                    "subnet_xyz": True},
}


def make_decoder(seed=0, kind="default", device="cuda", precision="bf16", prefix=""):
    """Product Decoder with the same synthetic state dict the oracle folds."""
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    dec = Decoder(latent_size=128, tool_latent_size=64, num_dims=3, do_code_regularization=True,
                  **SPECS["NetworkSpecs"], **SPECS["SubnetSpecs"])
    dec.load_state_dict(synth.make_state_dict(seed, kind, prefix=prefix))
    dec.precision = precision
    return dec.to(device).eval()


@functools.lru_cache(maxsize=None)
def oracle_net(seed=0, kind="default", dtype="float32"):
    return O.fold(synth.make_state_dict(seed, kind), dtype=getattr(torch, dtype))


def trained_code(seed=0):
    rows = np.load(os.path.join(GOLDEN, "latent_rows.npz"))
    return torch.from_numpy(np.concatenate([rows["c2000"][seed % 8], rows["c2000_tool"][seed % 8]]))


def code_for(seed, kind):
    return trained_code(seed) if kind == "trained_norm" else synth.make_code(seed)
