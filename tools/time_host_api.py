"""Host-facing calls of the reference API at 256^3: decode_samples (f64 points in, f64 values out), decode_shape."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200.core import reconstruct as R
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, os.environ.get("PREC", "bf16"))
d = int(os.environ.get("DIMS", "256"))
code = synth_data.trained_code(0)
pts = R.get_query_points((d, d, d), 0.1)
for name, fn in (("decode_samples", lambda: R.decode_samples(code, dec, pts, use_net=1, sigmoid=False)),
                 ("decode_shape", lambda: R.decode_shape(code, dec, (d, d, d), use_net=1, padding=0.1, sigmoid=False))):
    for _ in range(2):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(5):
        out = fn()
    print("%s %d^3: %.1f ms per call, out %s %s" % (name, d, 1e3 * (time.perf_counter() - t0) / 5, out.shape, out.dtype))
a = R.decode_samples(code, dec, pts, use_net=1, sigmoid=False).reshape(-1)
b = R.decode_shape(code, dec, (d, d, d), use_net=1, padding=0.1, sigmoid=False, reshape=False).reshape(-1)
print("decode_samples == decode_shape:", bool(np.array_equal(a, b)))
