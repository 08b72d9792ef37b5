"""Worker-side glue with the call surface of the reference's ``core/utils_multiprocessing.py:16-114``
(``load_network``, ``reconstruct_chunk``, ``mesh_chunk``: same names, keyword arguments, skip / overwrite rules
and error handling), one decoder and one CUDA context per worker process."""
import logging
import os

import numpy as np
import torch

from deepjoin_b200.mesh import Mesh
from . import errors

model_params_subdir = "ModelParameters"  # experiment layout, core/workspace.py:8
_log = logging.getLogger(__name__)


def file_is_old(path, overwrite_time=1636410000):
    """A result written before ``overwrite_time`` counts as missing (core/utils.py:6-18)."""
    return overwrite_time is not None and os.path.exists(path) and os.path.getmtime(path) < overwrite_time


def lossify_log_path(path):
    """``a/b.pth`` -> ``a/b__loss.pth`` (core/handler.py:617-619)."""
    stem, ext = os.path.splitext(path)
    return stem + "__loss" + ext


def load_network(decoder_kwargs, decoder_constructor, experiment_directory, checkpoint):
    """core/utils_multiprocessing.py:16-30.  The reference wraps the decoder in DataParallel only to match the
    checkpoint's ``module.`` key prefix; ``Decoder.load_state_dict`` here strips that prefix itself."""
    ckpt = os.path.join(experiment_directory, model_params_subdir, checkpoint + ".pth")
    state = torch.load(ckpt, map_location="cpu", weights_only=False)["model_state_dict"]
    decoder = decoder_constructor(**decoder_kwargs)
    decoder.load_state_dict(state)
    return decoder.cuda().eval()


def _todo(path, overwrite):
    return overwrite or not os.path.exists(path) or file_is_old(path)


def _each(chunk_inputs, chunk_paths, overwrite, callback, work):
    """The shared item loop of both workers: skip finished items, report progress.  ``work`` may return a callable
    with the host-only rest of the item (exporting a mesh): it runs on a background thread while the next item's GPU
    work is enqueued, at most one in flight; progress is reported when an item is really finished."""
    import threading

    def done():
        if callback is not None:
            callback()

    pending, pending_path = None, None
    for item, path in zip(chunk_inputs, chunk_paths):
        if pending is not None and path == pending_path:  # the same file again: let the first write land
            pending.join()
            done()
            pending = None
        rest = None
        if _todo(path, overwrite):
            rest = work(item, path)
        else:
            _log.debug("Skipping %s", path)
        if pending is not None:
            pending.join()
            done()
            pending = None
        if rest is None:
            done()
        else:
            pending, pending_path = threading.Thread(target=rest), path
            pending.start()
    if pending is not None:
        pending.join()
        done()


def reconstruct_chunk(chunk_inputs, chunk_paths, reconstruct_fn, network_kwargs, reconstruction_kwargs,
                      overwrite=False, callback=None, save=True):
    """core/utils_multiprocessing.py:33-71: latent code + loss log per item (``Codes/<idx>_<sig>.pth``,
    ``...__loss.npy``)."""
    decoder = load_network(**network_kwargs)

    def work(item, path):
        loss, latent = reconstruct_fn(decoder=decoder, **item, **reconstruction_kwargs)
        if not save:
            return
        try:
            _log.debug("Reconstruct chunker saving to: %s", path)
            torch.save(latent, path)
            np.save(lossify_log_path(path), loss)
        except PermissionError:
            _log.warning("Could not save: %s", path)

    _each(chunk_inputs, chunk_paths, overwrite, callback, work)


def mesh_chunk(chunk_inputs, chunk_paths, mesh_fn, network_kwargs, mesh_kwargs, overwrite=False, callback=None,
               save=True):
    """core/utils_multiprocessing.py:74-114: one mesh file per item; an IsosurfaceExtractionError becomes an
    empty mesh (:96-98), a PermissionError is logged and the item skipped."""
    decoder = load_network(**network_kwargs)

    def work(item, path):
        try:
            with torch.no_grad():
                mesh = mesh_fn(decoder=decoder, **item, **mesh_kwargs)
        except errors.IsosurfaceExtractionError:
            _log.debug("Isosurface extraction error on sample: %s", path)
            mesh = Mesh()
        except PermissionError:
            _log.warning("Could not save: %s", path)
            return
        if not save:
            return None

        def export():  # host-only (the library's PLY writer releases the GIL): overlaps the next item's query
            try:
                _log.debug("Mesh chunker saving to: %s", path)
                mesh.export(path)
            except PermissionError:
                _log.warning("Could not save: %s", path)

        return export

    _each(chunk_inputs, chunk_paths, overwrite, callback, work)
