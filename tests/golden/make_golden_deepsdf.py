"""Golden vectors of the reference's baseline decoder (networks/decoder_deepsdf.py), produced by importing
the reference's own file from /root/reference (build container only) -> tests/golden/deepsdf_forward.npz."""
import importlib
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synth_data  # noqa: E402

REF_PY = "/root/reference/deepjoin/python"
SPECS = os.path.join(REF_PY, "..", "experiments", "mugs_deepsdf", "specs.json")


def main():
    sys.path.insert(0, REF_PY)
    arch = importlib.import_module("networks.decoder_deepsdf")
    specs = json.load(open(SPECS))
    out = {}
    for kind in ("default", "trained_norm"):
        for seed in (0, 1):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                dec = arch.Decoder(specs["CodeLength"], **specs["NetworkSpecs"]).eval()
            dec.load_state_dict(synth_data.make_deepsdf_state_dict(seed, kind))
            code = synth_data.trained_code(seed)[:128] if kind == "trained_norm" else synth_data.make_code(seed)[:128]
            pts = synth_data.make_points(2048, 20 + seed)
            with torch.no_grad():
                ref = dec(torch.cat([code.reshape(1, -1).expand(pts.shape[0], -1), pts], 1))
            out["%s_%d_code" % (kind, seed)] = code.numpy()
            out["%s_%d_pts" % (kind, seed)] = pts.numpy()
            out["%s_%d_out" % (kind, seed)] = ref.numpy()
            print(kind, seed, "sdf range", float(ref.min()), float(ref.max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "deepsdf_forward.npz"), **out)


if __name__ == "__main__":
    main()
