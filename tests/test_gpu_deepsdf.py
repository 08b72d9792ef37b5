"""the DeepSDF baseline decoder and its grid query / meshing on the fused kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_oracle as O
from oracle import mc_oracle, synth
from tests import util

pytestmark = pytest.mark.gpu
G = util.GOLDEN
SPECS = dict(dims=[512] * 8, dropout=list(range(8)), dropout_prob=0.2, norm_layers=list(range(8)), latent_in=[4],
             xyz_in_all=False, use_tanh=False, latent_dropout=False, weight_norm=True)


def _decoder(seed, kind, precision):
    from deepjoin_b200.networks.decoder_deepsdf import Decoder
    dec = Decoder(128, **SPECS)
    dec.load_state_dict(synth.make_deepsdf_state_dict(seed, kind, prefix="module."))
    dec.precision = precision
    return dec.cuda().eval()


@pytest.mark.parametrize("kind,precision,tol", [("default", "fp32", 2e-5), ("trained_norm", "fp32", 3e-4),
                                                ("default", "bf16", 1e-3), ("trained_norm", "bf16", 0.3)])
@pytest.mark.parametrize("seed", [0, 1])
def test_forward_matches_the_reference_decoder(kind, precision, tol, seed):
    g = np.load(os.path.join(G, "deepsdf_forward.npz"))
    dec = _decoder(seed, kind, precision)
    code = torch.from_numpy(g["%s_%d_code" % (kind, seed)])
    pts = torch.from_numpy(g["%s_%d_pts" % (kind, seed)])
    out = dec(torch.cat([code.reshape(1, -1).expand(pts.shape[0], -1), pts], 1).cuda()).cpu().numpy()
    ref = g["%s_%d_out" % (kind, seed)]
    assert out.shape == ref.shape and np.abs(out - ref).max() <= tol
    if precision == "bf16" and kind == "trained_norm":
        assert np.abs(out - ref).mean() < 2e-2


def test_create_mesh_and_values2mesh():
    from deepjoin_b200.core import reconstruct as R, utils_deepsdf as U
    from deepjoin_b200.mesh import Mesh
    dec = _decoder(0, "trained_norm", "fp32")
    latent = synth.trained_code(0)[:128]
    N = 40
    mesh = U.create_mesh(dec, latent.reshape(1, -1), N=N, padding=0.1)
    vol = R.decode_shape_device(dec.code192(latent), dec, (N, N, N), use_net=0, padding=0.1, sigmoid=False).reshape(N, N, N)
    # the grid itself against the oracle restatement of the reference decoder
    pts = torch.from_numpy(R.get_query_points((N, N, N), 0.1)).float()
    ref = O.deepsdf_forward(synth.make_deepsdf_state_dict(0, "trained_norm"), latent, pts).numpy().reshape(N, N, N)
    assert np.abs(vol.cpu().numpy() - ref).max() < 3e-4
    vo, fo = mc_oracle.mesh_from_volume(vol.cpu().numpy(), 0.0, (N, N, N), 0.1, "ascent")
    want = Mesh(vo, fo)
    assert np.array_equal(mesh.faces, want.faces) and np.array_equal(mesh.vertices, want.vertices)
    again = U.values2mesh(vol.cpu().numpy(), [-(0.5 + 0.1)] * 3, (1 + 0.2) / N, gradient_direction="ascent", flipxy=True)
    assert np.array_equal(again.faces, want.faces) and np.allclose(again.vertices, want.vertices, rtol=0, atol=1e-12)
    sdf = U.decode_sdf(dec, latent.reshape(1, -1), pts[:100])
    assert np.abs(sdf.cpu().numpy().reshape(-1) - ref.reshape(-1)[:100]).max() < 3e-4
