"""Drop-in for the grid-query / meshing part of the reference's ``core/utils_deepsdf.py`` (the DeepSDF
baseline): ``decode_sdf`` (:11-20), ``create_mesh`` (:123-175), ``values2mesh`` (:178-208).  Same grid
(np.meshgrid 'xy' order), same marching-cubes call and vertex post-processing as DeepJoin's create_mesh,
so it runs through the same C-ABI entry points.  Line numbers cite the reference file."""
import logging

import numpy as np
import torch

from deepjoin_b200.mesh import Mesh
from . import errors
from . import reconstruct as _R


def decode_sdf(decoder, latent_vector, queries):
    """:11-20.  queries [N,3] tensor; latent_vector [1,L] / [L] (None: queries already carry the code)."""
    if latent_vector is None:
        return decoder(queries.cuda())
    return decoder.query(latent_vector.reshape(-1), queries.cuda())


def create_mesh(decoder, latent_vec, N=256, max_batch=32 ** 3, padding=0.1, save_values=None):
    """:123-175: N^3 grid query (max_batch accepted and ignored: chunking is internal), marching cubes at
    level 0 with gradient_direction='ascent', x/y flipped, voxel origin -(0.5+padding)."""
    decoder.eval()
    dims = (N, N, N)
    code = decoder.code192(latent_vec if isinstance(latent_vec, torch.Tensor) else torch.as_tensor(latent_vec))
    if save_values is not None:
        logging.debug("Saving values to {}".format(save_values))
        vals = _R.decode_shape_device(code, decoder, dims, use_net=0, padding=padding, sigmoid=False)
        np.save(save_values, vals.reshape(dims).cpu().numpy())
    return _R.create_mesh(code, decoder, 0, sigmoid=False, level=0.0, gradient_direction="ascent", dims=dims, padding=padding)


def values2mesh(sdf_values, voxel_grid_origin, voxel_size, gradient_direction="descent", level=0.0, flipxy=False):
    """:178-208 on a host (or device) volume: marching cubes with spacing voxel_size, optional x/y flip,
    translation by voxel_grid_origin, trimesh-style vertex merge."""
    vol = sdf_values if isinstance(sdf_values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(sdf_values, np.float32))
    if not vol.is_cuda:
        vol = vol.cuda()
    try:
        verts, faces = _R.marching_cubes(vol, level, gradient_direction)
    except ValueError:
        raise errors.IsosurfaceExtractionError
    v = verts.to(torch.float64) * float(voxel_size)
    if flipxy:
        v = torch.stack([v[:, 1], v[:, 0], v[:, 2]], dim=1)
    v = v + torch.tensor([float(o) for o in voxel_grid_origin], dtype=torch.float64, device=v.device)
    v, f = _R.merge_vertices(v, faces)
    return Mesh(v.cpu().numpy(), f.cpu().numpy(), process=False)
