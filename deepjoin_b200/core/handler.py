"""File naming and load / save of the inference path's artefacts (reference core/handler.py): the batch drivers
address every code, loss log, value grid and mesh through a ``ReconstructionHandler`` (reconstruct.py:569-725), so a
drop-in must produce the same paths:

    Reconstructions/<name>@<ckpt>/Codes/<idx>_<lat_signiture>.pth            path_code   (handler.py:787-796)
    Reconstructions/<name>@<ckpt>/Codes/<idx>_<lat_signiture>__loss.pth      path_loss   (:806-807, :617-619)
    Reconstructions/<name>@<ckpt>/Meshes/<idx>_<shape>_<signiture>.ply       path_mesh   (:809-818)
    Reconstructions/<name>@<ckpt>/Meshes/<idx>_<shape>_<signiture>_values.npy path_values (:845-860)
    Reconstructions/<name>@<ckpt>/Codes/<idx>_<shape>_<gt>_<metric>_<signiture>.npy  path_eval (:938-956)

with signiture = "".join(str(s) + "_" for s in signiture) (:664-665).  Rendering, summaries, interpolation, the
dataset handler and the knit/pickle plumbing of that 1,500-line file are outside the hot path and not provided."""
import json
import logging
import os
import pickle

import numpy as np
import torch

from . import errors
from . import workspace as ws
from deepjoin_b200.mesh import Mesh, load_obj, load_ply


def iterify_path(path, i):
    """a/b.ext -> a/b__iter-<i>.ext (handler.py:612-614)."""
    stem, ext = os.path.splitext(path)
    return "{}__iter-{}{}".format(stem, i, ext)


def lossify_log_path(path):
    """a/b.ext -> a/b__loss.ext (handler.py:617-619)."""
    stem, ext = os.path.splitext(path)
    return stem + "__loss" + ext


_READERS = {
    ".pkl": lambda f: pickle.load(open(f, "rb")),
    ".json": lambda f: json.load(open(f, "r")),
    ".npy": lambda f: np.load(f, allow_pickle=True),
    ".npz": lambda f: np.load(f, allow_pickle=True),
    ".pth": lambda f: torch.load(f, map_location=torch.device("cpu"), weights_only=False),
    ".ply": load_ply,
    ".obj": load_obj,
}
_WRITERS = {
    ".pkl": lambda f, d: pickle.dump(d, open(f, "wb"), pickle.HIGHEST_PROTOCOL),
    ".json": lambda f, d: json.dump(d, open(f, "w")),
    ".npy": lambda f, d: np.save(f, d),
    ".npz": lambda f, d: np.savez(f, d),
    ".pth": lambda f, d: torch.save(d, f),
    ".ply": lambda f, d: d.export(f),
    ".obj": lambda f, d: d.export(f),
}


def loader(f_in):
    """handler.py:36-78 for the file types the path touches (images / gifs are rendering outputs)."""
    ext = os.path.splitext(f_in)[-1]
    logging.debug("Loading file {}".format(f_in))
    if ext not in _READERS:
        raise RuntimeError("Loader: Unhandled file type: {}".format(f_in))
    return _READERS[ext](f_in)


def saver(f_out, data, all_access=True, **kwargs):
    """handler.py:81-126: files are created world-accessible first, as the reference does for shared result trees."""
    ext = os.path.splitext(f_out)[-1]
    logging.debug("Saving file {}".format(f_out))
    if ext not in _WRITERS:
        raise RuntimeError("Saver: Unhandled file type: {}".format(f_out))
    if all_access:
        os.close(os.open(f_out, os.O_CREAT | os.O_WRONLY, 0o777))
    return _WRITERS[ext](f_out, data)


class ReconstructionHandler:
    """Paths, cache and load / save of codes, loss logs, value grids and meshes of one reconstruction run
    (handler.py:622-1040).  ``signiture`` / ``lat_signiture``: lists of run-identifying values (lr, iterations, ...)."""

    def __init__(self, experiment_directory, signiture, name, checkpoint, lat_signiture=None, use_occ=False, decoder=None,
                 dims=(256, 256, 256), overwrite=False):
        assert os.path.exists(experiment_directory), "{}".format(experiment_directory)
        join = lambda sig: "".join(str(s) + "_" for s in sig)
        self._experiment_directory = os.path.abspath(experiment_directory)
        self._signiture = join(signiture)
        self._lat_signiture = join(signiture if lat_signiture is None else lat_signiture)
        self._name, self._checkpoint = name, checkpoint
        self._decoder, self._dims, self._overwrite, self._use_occ = decoder, dims, overwrite, use_occ
        self._level = 0.5 if use_occ else 0.0
        self._cache = {}

    overwrite = property(lambda self: self._overwrite)
    signiture = property(lambda self: self._signiture)
    name = property(lambda self: self._name)
    root = property(lambda self: self._experiment_directory)
    experiment_name = property(lambda self: self._experiment_directory.split("/")[-1])

    # ---- paths
    def _dir(self, getter, create):
        return getter(self._experiment_directory, name=self._name, checkpoint=self._checkpoint, create=create)

    def dir_codes(self, create=False):
        return self._dir(ws.get_reconstructions_codes_dir, create)

    def dir_meshes(self, create=False):
        return self._dir(ws.get_reconstructions_meshes_dir, create)

    def path_code(self, idx, create=False):
        return os.path.join(self.dir_codes(create), "{}_{}.pth".format(idx, self._lat_signiture))

    def path_loss(self, idx, create=False):
        return lossify_log_path(self.path_code(idx, create))

    def path_mesh(self, idx, shape_idx, create=False):
        return os.path.join(self.dir_meshes(create), "{}_{}_{}.ply".format(idx, shape_idx, self._signiture))

    def path_values(self, idx, shape_idx, create=False):
        return os.path.join(self.dir_meshes(create), "{}_{}_{}_values.npy".format(idx, shape_idx, self._signiture))

    def path_eval(self, idx, shape_idx, gt_shape_idx, metric, create=False):
        return os.path.join(self.dir_codes(create), "{}_{}_{}_{}_{}.npy".format(idx, shape_idx, gt_shape_idx, metric, self._signiture))

    # ---- cache + files
    def load(self, f_in, skip_cache=False):
        if f_in in self._cache and not skip_cache:
            return self._cache[f_in]
        data = loader(f_in)
        if not skip_cache:
            self._cache[f_in] = data
        return data

    def delete_from_cache(self, path):
        self._cache.pop(path, None)

    def _set(self, path, data, save, cache=True):
        if save:
            saver(path, data)
        if cache:
            self._cache[path] = data

    def set_code(self, idx, code, save=False, cache=True):
        self._set(self.path_code(idx, create=True), code, save, cache)

    def get_code(self, idx):
        """FileNotFoundError if the code was never produced (handler.py:985-989)."""
        return self.load(self.path_code(idx, create=True))

    def set_values(self, idx, shape_idx, values, save=False, cache=True):
        self._set(self.path_values(idx, shape_idx, create=True), values, save, cache)

    def get_values(self, idx, shape_idx):
        return self.load(self.path_values(idx, shape_idx, create=True))

    def set_mesh(self, mesh, idx, shape_idx, save=False):
        self._set(self.path_mesh(idx, shape_idx, create=True), mesh, save)

    def get_mesh(self, idx, shape_idx):
        """The cached or stored mesh; an unreadable or empty one is an IsosurfaceExtractionError (handler.py:1019-1038)."""
        path = self.path_mesh(idx, shape_idx, create=True)
        if self._overwrite:
            return None
        try:
            mesh = self.load(path)
        except (ValueError, IndexError, FileNotFoundError):  # trimesh reports a missing file as ValueError (:1032-1034)
            raise errors.IsosurfaceExtractionError
        if mesh is None or not isinstance(mesh, Mesh) or mesh.vertices.shape[0] == 0:
            raise errors.IsosurfaceExtractionError
        return mesh
