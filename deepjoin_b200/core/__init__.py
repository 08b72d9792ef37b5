"""Host-side mirror of the reference's ``core`` package, restricted to the inference hot path
(SURVEY.md §8): errors, reconstruct, data (sampler), utils_multiprocessing (worker glue), utils_deepsdf (baseline decoder's
grid query / meshing), workspace / handler (where results are named and stored, §8 f3)."""
from . import errors  # noqa: F401  (first: _lib imports it)
from .errors import *  # noqa: F401,F403
from . import data, metrics, reconstruct, utils_multiprocessing, utils_deepsdf, workspace, handler  # noqa: F401,E402
from .reconstruct import (create_mesh, decode_samples, decode_shape, get_query_points,  # noqa: F401,E402
                          tensor_sdf_to_occ, tensor_sdf_to_udf, tensor_udf_to_sdf)
from .data import select_samples, clamp_samples  # noqa: F401,E402
from .utils_multiprocessing import load_network, reconstruct_chunk, mesh_chunk  # noqa: F401,E402
from .handler import loader, saver, ReconstructionHandler  # noqa: F401,E402
