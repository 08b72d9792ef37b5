"""Throughput of the batch meshing worker (core.utils_multiprocessing.mesh_chunk): 256^3 meshes written as PLY."""
import os, sys, time, tempfile, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import SPECS
from deepjoin_b200 import core
from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
d = tempfile.mkdtemp()
os.makedirs(os.path.join(d, "ModelParameters"))
torch.save({"epoch": 2000, "model_state_dict": synth_data.make_state_dict(0, "trained_norm")}, os.path.join(d, "ModelParameters", "2000.pth"))
net = dict(decoder_kwargs=dict(latent_size=128, tool_latent_size=64, num_dims=3, do_code_regularization=True, **SPECS["NetworkSpecs"], **SPECS["SubnetSpecs"]),
           decoder_constructor=Decoder, experiment_directory=d, checkpoint="2000")
kw = dict(sigmoid=False, level=0.0, gradient_direction="ascent", dims=[256] * 3, padding=0.1)
vec = synth_data.trained_code(0)
for rep in range(2):
    n = 8
    items = [dict(vec=vec, use_net=1 + (i % 2)) for i in range(n)]
    paths = [os.path.join(d, "r%d_%d.ply" % (rep, i)) for i in range(n)]
    t0 = time.perf_counter()
    core.mesh_chunk(items, paths, core.reconstruct.create_mesh, net, kw)
    dt = time.perf_counter() - t0
    print("mesh_chunk: %d meshes at 256^3 in %.1f ms = %.1f ms per mesh (incl. loading the network), files %.1f MB" % (n, 1e3 * dt, 1e3 * dt / n, sum(os.path.getsize(p) for p in paths) / 1e6))
