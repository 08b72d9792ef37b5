"""Where create_mesh's host-side time goes: count call (query + MC + merge), fetch (D2H), mesh object."""
import os, sys, time, ctypes, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import _lib
from deepjoin_b200.core import reconstruct as R
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, "bf16")
d = int(os.environ.get("DIMS", "256"))
code = synth_data.trained_code(0)
for _ in range(3):
    m = R.create_mesh(code, dec, 1, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(d, d, d), padding=0.1)
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for i in range(n):
    m = R.create_mesh(code, dec, 1, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(d, d, d), padding=0.1)
print("create_mesh %.3f ms per call (%d vertices, %d faces, faces %s)" % (1e3 * (time.perf_counter() - t0) / n, len(m.vertices), len(m.faces), m.faces.dtype))
