"""Where the reference keeps things inside an experiment directory (reference core/workspace.py:8-21,113-203),
restricted to what the inference path reads and writes: specs.json, ModelParameters/<ckpt>.pth and the
Reconstructions/<name>@<ckpt>/{Codes,Meshes,Stats} tree the batch drivers fill (reconstruct.py:569-725)."""
import json
import logging
import os

import torch

# directory / file names inside an experiment (module attributes <key>_subdir, <key>_filename as in the reference)
_NAMES = {
    "model_params_subdir": "ModelParameters", "latent_codes_subdir": "LatentCodes",
    "reconstructions_subdir": "Reconstructions", "reconstruction_meshes_subdir": "Meshes",
    "reconstruction_codes_subdir": "Codes", "reconstruction_stats_subdir": "Stats",
    "logs_filename": "Logs.pth", "specifications_filename": "specs.json",
}
globals().update(_NAMES)


def load_experiment_specifications(experiment_directory):
    """workspace.py:24-32."""
    filename = os.path.join(experiment_directory, _NAMES["specifications_filename"])
    if not os.path.isfile(filename):
        raise Exception("The experiment directory ({}) does not include specifications file ".format(experiment_directory)
                        + '"specs.json"')
    with open(filename) as f:
        return json.load(f)


def load_model_parameters(experiment_directory, checkpoint, decoder):
    """workspace.py:34-42: state dict of ModelParameters/<checkpoint>.pth into ``decoder``; returns the epoch."""
    filename = os.path.join(experiment_directory, _NAMES["model_params_subdir"], checkpoint + ".pth")
    if not os.path.isfile(filename):
        raise Exception('model state dict "{}" does not exist'.format(filename))
    data = torch.load(filename, map_location="cpu", weights_only=False)
    decoder.load_state_dict(data["model_state_dict"])
    return data["epoch"]


def build_decoder(experiment_directory, experiment_specs):
    """workspace.py:45-55: the decoder class is found by name (the drop-in's plug-in point)."""
    arch = __import__("networks." + experiment_specs["NetworkArch"], fromlist=["Decoder"])
    return arch.Decoder(experiment_specs["CodeLength"], **experiment_specs["NetworkSpecs"]).cuda()


def _made(path, create):
    if create and not os.path.isdir(path):
        logging.debug("Creating directory: {}".format(path))
        os.makedirs(path)
    return path


def get_model_params_dir(experiment_dir, create_if_nonexistent=False):
    return _made(os.path.join(experiment_dir, _NAMES["model_params_subdir"]), create_if_nonexistent)


def get_reconstructions_dir(experiment_directory, name=None, checkpoint="latest", create=False):
    """Reconstructions/<checkpoint> or Reconstructions/<name>@<checkpoint> (workspace.py:113-131)."""
    leaf = str(checkpoint) if name is None else name + "@" + str(checkpoint)
    return _made(os.path.join(experiment_directory, _NAMES["reconstructions_subdir"], leaf), create)


def _reconstructions_sub(sub):
    def get(experiment_directory, name=None, checkpoint="latest", create=False):
        return _made(os.path.join(get_reconstructions_dir(experiment_directory, name, checkpoint, create), sub), create)
    return get


get_reconstructions_codes_dir = _reconstructions_sub(_NAMES["reconstruction_codes_subdir"])     # workspace.py:134-149
get_reconstructions_stats_dir = _reconstructions_sub(_NAMES["reconstruction_stats_subdir"])     # workspace.py:152-167
get_reconstructions_meshes_dir = _reconstructions_sub(_NAMES["reconstruction_meshes_subdir"])   # workspace.py:188-203
