"""Marching cubes alone on the 256^3 bench volume (for ncu launch lists / captures of the MC kernels)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200.core import reconstruct as R
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, os.environ.get("PREC", "bf16"))
d = int(os.environ.get("DIMS", "256"))
vol = R.decode_shape_device(synth_data.trained_code(0).to(dev), dec, (d, d, d), use_net=1, padding=0.1, sigmoid=False).reshape(d, d, d)
for _ in range(4):
    v, f = R.marching_cubes(vol, 0.0, "ascent")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    v, f = R.marching_cubes(vol, 0.0, "ascent")
e1.record(); torch.cuda.synchronize()
print("marching cubes %d^3: %.3f ms per call, %d vertices, %d faces" % (d, e0.elapsed_time(e1) / 10, v.shape[0], f.shape[0]))
