"""The reference's pool workers (core/utils_multiprocessing.py:16-114) on the product path: checkpoint loading
with the DataParallel ``module.`` prefix, per-item files, skip / overwrite rules, empty mesh on an
IsosurfaceExtractionError."""
import os

import numpy as np
import pytest
import torch

from oracle import synth
from tests import util

pytestmark = pytest.mark.gpu


def _experiment(tmp_path, kind):
    os.makedirs(tmp_path / "ModelParameters")
    torch.save({"epoch": 2000, "model_state_dict": synth.make_state_dict(0, kind, prefix="module.")},
               str(tmp_path / "ModelParameters" / "2000.pth"))
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    kwargs = dict(latent_size=128, tool_latent_size=64, num_dims=3, do_code_regularization=True,
                  **util.SPECS["NetworkSpecs"], **util.SPECS["SubnetSpecs"])
    return dict(decoder_kwargs=kwargs, decoder_constructor=Decoder, experiment_directory=str(tmp_path), checkpoint="2000")


def test_reconstruct_chunk_and_mesh_chunk(tmp_path):
    from deepjoin_b200 import core
    from deepjoin_b200.mesh import load_ply
    from deepjoin_b200.reconstruct import reconstruct
    net = _experiment(tmp_path, "trained_norm")
    xyz, sdf, nrm = synth.make_sample_set(20000, seed=4)
    items = [dict(test_sdf=[xyz, sdf, nrm])] * 2
    code_paths = [str(tmp_path / ("%d_code.pth" % i)) for i in range(2)]
    calls = []
    rec_kwargs = dict(num_iterations=20, latent_size=128, break_latent_size=64, stat=0.01, lambda1=1.0, lambda2=0.0, lambda3=1.0,
                      clamp_dist=0.1, num_samples=1500, lr=5e-3, l2reg=True, code_reg_lambda=1e-4, sampling_method="device")
    np.random.seed(0)
    torch.manual_seed(0)
    core.reconstruct_chunk(items, code_paths, reconstruct, net, rec_kwargs, callback=lambda: calls.append(1))
    assert len(calls) == 2
    for p in code_paths:
        code = torch.load(p, weights_only=False)
        assert tuple(code.shape) == (1, 192)
        loss = np.load(core.utils_multiprocessing.lossify_log_path(p).replace(".pth", ".pth.npy"), allow_pickle=True).item()
        assert loss["epoch"] == [0, 10] and loss["loss"][-1] < loss["loss"][0]
    stamp = os.path.getmtime(code_paths[0])
    core.reconstruct_chunk(items, code_paths, reconstruct, net, rec_kwargs)           # finished items are skipped
    assert os.path.getmtime(code_paths[0]) == stamp
    # meshes from the shipped latent row (a code with a surface for these synthetic weights)
    mesh_items = [dict(vec=util.trained_code(0), use_net=u) for u in (0, 1)]
    mesh_paths = [str(tmp_path / ("0_%d_mesh.ply" % u)) for u in (0, 1)]
    mesh_kwargs = dict(sigmoid=False, level=0.0, gradient_direction="ascent", dims=[40] * 3, batch_size=2 ** 14, padding=0.1)
    core.mesh_chunk(mesh_items, mesh_paths, core.reconstruct.create_mesh, net, mesh_kwargs)
    for p in mesh_paths:
        assert len(load_ply(p).faces) > 0


def test_mesh_chunk_writes_an_empty_mesh_when_there_is_no_surface(tmp_path):
    from deepjoin_b200 import core
    from deepjoin_b200.mesh import load_ply
    net = _experiment(tmp_path, "default")   # degenerate field: no zero level set
    path = str(tmp_path / "none.ply")
    core.mesh_chunk([dict(vec=util.code_for(0, "default"), use_net=0)], [path], core.reconstruct.create_mesh, net,
                    dict(sigmoid=False, level=0.0, gradient_direction="ascent", dims=[16] * 3, padding=0.1))
    assert load_ply(path).is_empty


def test_batch_driver_flow_through_the_reconstruction_handler(tmp_path):
    """What reconstruct.py:569-725 does: the handler names every artefact, the pool workers fill the tree, the handler
    reads it back (codes from Codes/<idx>_<sig>.pth, meshes from Meshes/<idx>_<shape>_<sig>.ply)."""
    from deepjoin_b200 import core
    from deepjoin_b200.reconstruct import reconstruct
    net = _experiment(tmp_path, "trained_norm")
    h = core.handler.ReconstructionHandler(str(tmp_path), signiture=[20, 0.005, 1500], name="ours", checkpoint="2000")
    xyz, sdf, nrm = synth.make_sample_set(20000, seed=4)
    idxs = [3, 5]
    rec_kwargs = dict(num_iterations=20, latent_size=128, break_latent_size=64, stat=0.01, lambda1=1.0, lambda2=0.0, lambda3=1.0,
                      clamp_dist=0.1, num_samples=1500, lr=5e-3, l2reg=True, code_reg_lambda=1e-4, sampling_method="device")
    np.random.seed(0)
    torch.manual_seed(0)
    core.reconstruct_chunk([dict(test_sdf=[xyz, sdf, nrm])] * 2, [h.path_code(i, create=True) for i in idxs], reconstruct, net, rec_kwargs)
    assert sorted(os.listdir(h.dir_codes())) == sorted(
        ["3_20_0.005_1500_.pth", "3_20_0.005_1500___loss.pth.npy", "5_20_0.005_1500_.pth", "5_20_0.005_1500___loss.pth.npy"])
    codes = [h.get_code(i) for i in idxs]
    assert all(tuple(c.shape) == (1, 192) for c in codes)
    # restoration / fractured / complete shapes of one code (the shipped latent row: a code with a surface for these weights)
    vec = util.trained_code(0)
    core.mesh_chunk([dict(vec=vec, use_net=u) for u in (0, 1, 2)], [h.path_mesh(3, u, create=True) for u in (0, 1, 2)],
                    core.reconstruct.create_mesh, net,
                    dict(sigmoid=False, level=0.0, gradient_direction="ascent", dims=[40] * 3, padding=0.1))
    direct = core.reconstruct.create_mesh(vec, core.load_network(**net), 1, sigmoid=False, level=0.0, gradient_direction="ascent",
                                          dims=[40] * 3, padding=0.1)
    back = h.get_mesh(3, 1)
    assert np.array_equal(back.faces, direct.faces) and np.allclose(back.vertices, direct.vertices, atol=1e-6)  # PLY stores f32
    with pytest.raises(core.errors.IsosurfaceExtractionError):
        h.get_mesh(5, 1)  # never meshed
