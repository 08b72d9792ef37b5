"""How does tcgen05.mma accumulate in fp32?  bf16-exact operands (exact products), compare with the fp64 sum."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the dj_debug_* entry points live in the diagnostics build only (python -m deepjoin_b200.build --debug)
os.environ.setdefault("DEEPJOIN_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deepjoin_b200", "lib", "libdeepjoin_b200_debug.so"))
import numpy as np
import torch
from deepjoin_b200 import _lib

def bf16_exact(x):
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()

def main():
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for name, gen in (("positive", lambda s: np.exp(rng.normal(0, 1.5, s))),
                      ("mostly pos", lambda s: np.exp(rng.normal(0, 1.5, s)) * np.where(rng.uniform(0, 1, s) < 0.15, -1.0, 1.0))):
        for K in (64, 128, 256):
            A = bf16_exact(gen((128, K)).astype(np.float32)); W = bf16_exact(gen((128, K)).astype(np.float32))
            D = np.zeros((128, 128), np.float32)
            rc = L.dj_debug_umma_probe(A.ctypes.data, W.ctypes.data, K, 1, 0, 0, 0, D.ctypes.data)
            assert rc == 0, _lib.last_error()
            S = A.astype(np.float64) @ W.astype(np.float64).T
            ulp = np.spacing(np.abs(S).astype(np.float32)).astype(np.float64)
            e = (D.astype(np.float64) - S) / ulp
            rn = (S.astype(np.float32).astype(np.float64) - S) / ulp
            # error signed toward zero (< 0: magnitude lost)
            tz = e * np.sign(S)
            print("%-10s K=%3d (%2d MMAs): err/ulp mean %+.3f  toward-zero mean %+.3f  rms %.3f  max|.| %.2f   (fp32 RN of the exact sum: rms %.3f)"
                  % (name, K, K // 16, e.mean(), tz.mean(), np.sqrt((e ** 2).mean()), np.abs(e).max(), np.sqrt((rn ** 2).mean())))

if __name__ == "__main__":
    main()
