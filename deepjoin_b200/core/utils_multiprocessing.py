"""Worker-side glue of the reference's ``core/utils_multiprocessing.py:16-114`` with the same
names and keyword arguments: one decoder (and CUDA context) per worker process."""
import logging
import os

import numpy as np
import torch

from deepjoin_b200.mesh import Mesh
from . import errors

model_params_subdir = "ModelParameters"  # core/workspace.py:8


def file_is_old(path, overwrite_time=1636410000):
    """core/utils.py:6-18."""
    if overwrite_time is None:
        return False
    return os.path.exists(path) and os.path.getmtime(path) < overwrite_time


def lossify_log_path(path):
    """core/handler.py:617-619."""
    p, e = os.path.splitext(path)
    return p + "__loss" + e


def load_network(decoder_kwargs, decoder_constructor, experiment_directory, checkpoint):
    """core/utils_multiprocessing.py:16-30.  The checkpoint's ``module.`` prefix (DataParallel)
    is accepted by Decoder.load_state_dict directly."""
    decoder = decoder_constructor(**decoder_kwargs)
    saved_model_state = torch.load(
        os.path.join(experiment_directory, model_params_subdir, checkpoint + ".pth"),
        map_location="cpu", weights_only=False)
    decoder.load_state_dict(saved_model_state["model_state_dict"])
    decoder = decoder.cuda()
    decoder.eval()
    return decoder


def reconstruct_chunk(chunk_inputs, chunk_paths, reconstruct_fn, network_kwargs, reconstruction_kwargs,
                      overwrite=False, callback=None, save=True):
    """core/utils_multiprocessing.py:33-71."""
    decoder = load_network(**network_kwargs)
    decoder.eval()
    for input, path in zip(chunk_inputs, chunk_paths):
        if overwrite or not os.path.exists(path) or file_is_old(path):
            loss, latent = reconstruct_fn(decoder=decoder, **input, **reconstruction_kwargs)
            try:
                if save:
                    logging.debug("Reconstruct chunker saving to: {}".format(path))
                    torch.save(latent, path)
                    np.save(lossify_log_path(path), loss)
            except PermissionError:
                logging.warning("Could not save: {}".format(path))
        else:
            logging.debug("Skipping {}".format(path))
        if callback is not None:
            callback()


def mesh_chunk(chunk_inputs, chunk_paths, mesh_fn, network_kwargs, mesh_kwargs, overwrite=False, callback=None,
               save=True):
    """core/utils_multiprocessing.py:74-114: IsosurfaceExtractionError -> empty mesh."""
    decoder = load_network(**network_kwargs)
    decoder.eval()
    for input, path in zip(chunk_inputs, chunk_paths):
        if overwrite or not os.path.exists(path) or file_is_old(path):
            try:
                with torch.no_grad():
                    mesh = mesh_fn(decoder=decoder, **input, **mesh_kwargs)
            except errors.IsosurfaceExtractionError:
                mesh = Mesh()
                logging.debug("Isosurface extraction error on sample: {}".format(path))
            except PermissionError:
                mesh = None
                logging.warning("Could not save: {}".format(path))
            try:
                if save and mesh is not None:
                    logging.debug("Mesh chunker saving to: {}".format(path))
                    mesh.export(path)
            except PermissionError:
                logging.warning("Could not save: {}".format(path))
        else:
            logging.debug("Skipping {}".format(path))
        if callback is not None:
            callback()
