"""Diagnostics for the three-pass split-operand forward (debug build): per-layer activation checksums
(DEEPJOIN_B200_DEBUG bit 6) of the three-pass program next to the four-pass one (DEEPJOIN_B200_X2_FOURPASS=1), the
error against the fp64 oracle, and structured first hidden layers (`ident`, `blocks`) that locate a wrong tile."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DEEPJOIN_B200_LIB", os.path.join(ROOT, "deepjoin_b200", "lib", "libdeepjoin_b200_debug.so"))
import numpy as np
import torch
from deepjoin_b200 import _lib
from oracle import decoder_oracle as O, synth
from tests import util


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def main():
    torch.cuda.set_device(0)
    seed = 0
    net64 = util.oracle_net(seed, "trained_norm", "float64")
    code = util.trained_code(seed)
    npts = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    pts = synth.make_points(npts, seed)
    dec = util.make_decoder(seed, "trained_norm", precision="fp16x2")
    pack = dec.pack(torch.device("cuda", 0))
    sums = {}
    for four in (None, "1"):
        setenv(DEEPJOIN_B200_DEBUG="64", DEEPJOIN_B200_X2_FOURPASS=four)
        out = dec.query(code.cuda(), pts.cuda(), use_net=1)
        torch.cuda.synchronize()
        buf = np.zeros(64, np.uint64)
        _lib.check(_lib.lib().dj_debug_read_prof(pack.handle, buf.ctypes.data, buf.size))
        sums[four] = buf.view(np.float64).reshape(4, 16)[:2]
    for net in range(2):
        for l in range(10):
            a, b = sums[None][net, l], sums["1"][net, l]
            if a or b:
                print("net %d layer %d: three-pass checksum %.6e  four-pass %.6e  rel diff %.2e" % (net, l, a, b, abs(a - b) / max(abs(b), 1e-30)))
    setenv(DEEPJOIN_B200_X2_FOURPASS=None)
    for use_net in (1,):
        ref = O.forward(net64, code.double(), pts.double(), use_net).numpy()
        for dbg in (None,):
            setenv(DEEPJOIN_B200_DEBUG=dbg)
            out = dec.query(code.cuda(), pts.cuda(), use_net=use_net).cpu().numpy().astype(np.float64)
            err = np.abs(out - ref).reshape(len(out), -1).max(1)
            bad = np.nonzero(err > 1e-2)[0]
            print("use_net %d debug %-3s: max %.3e mean %.3e  rows with err > 1e-2: %d of %d, first %s" % (use_net, dbg or "0", err.max(), err.mean(),
                  len(bad), len(err), bad[:12]), flush=True)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("ident", "blocks")):
    main()


def ident():
    """lin1 = diag(mask): layer-1 output column i = leaky(mask_i * a0_i); per 64-column group of K compare the layer-1
    checksum of the three-pass program with the four-pass one."""
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    torch.cuda.set_device(0)
    seed = 0
    code = util.trained_code(seed)
    pts = synth.make_points(int(sys.argv[2]) if len(sys.argv) > 2 else 128, seed)
    for m in range(8):
        sd = synth.make_state_dict(seed, "trained_norm")
        sd["lin1.weight_v"] = torch.eye(512)
        g = torch.zeros(512, 1)
        g[64 * m:64 * m + 64] = 1.0
        sd["lin1.weight_g"] = g
        sd["lin1.bias"] = torch.zeros(512)
        dec = Decoder(latent_size=128, tool_latent_size=64, num_dims=3, do_code_regularization=True,
                      **util.SPECS["NetworkSpecs"], **util.SPECS["SubnetSpecs"])
        dec.load_state_dict(sd)
        dec.precision = "fp16x2"
        dec = dec.cuda().eval()
        pack = dec.pack(torch.device("cuda", 0))
        res = {}
        for four in (None, "1"):
            for dbg in ("64",):
                setenv(DEEPJOIN_B200_DEBUG=dbg, DEEPJOIN_B200_X2_FOURPASS=four)
                dec.query(code.cuda(), pts.cuda(), use_net=1)
                torch.cuda.synchronize()
                buf = np.zeros(64, np.uint64)
                _lib.check(_lib.lib().dj_debug_read_prof(pack.handle, buf.ctypes.data, buf.size))
                res[(four, dbg)] = buf.view(np.float64).reshape(4, 16)[[0, 2], :3].copy()
        print("K columns [%3d, %3d) (range %d, k-chunk %d): layer-1 checksum three-pass %.6e four-pass %.6e | column-weighted: %.8e vs %.8e | layer 0 weighted: %.8e vs %.8e"
              % (64 * m, 64 * m + 64, m // 2, m % 2, res[(None, "64")][0, 1], res[("1", "64")][0, 1], res[(None, "64")][1, 1], res[("1", "64")][1, 1],
                 res[(None, "64")][1, 0], res[("1", "64")][1, 0]), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ident":
    ident()


def blocks():
    """lin1 = ones in one [64 x 64] block (output columns of group ng, K columns of group kg): which (ng, kg) tiles does
    the three-pass program get wrong?"""
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    torch.cuda.set_device(0)
    seed = 0
    code = util.trained_code(seed)
    pts = synth.make_points(128, seed)
    print("rows: output-column group ng, columns: K group kg; entry = relative difference of the layer-1 weighted checksum")
    for ng in range(8):
        line = []
        for kg in range(8):
            sd = synth.make_state_dict(seed, "trained_norm")
            v = torch.ones(512, 512) * 1e-3
            v[:, 64 * kg:64 * kg + 64] = 1.0
            g = torch.zeros(512, 1)
            g[64 * ng:64 * ng + 64] = float(torch.linalg.norm(v[0]))
            sd["lin1.weight_v"] = v
            sd["lin1.weight_g"] = g
            sd["lin1.bias"] = torch.zeros(512)
            dec = Decoder(latent_size=128, tool_latent_size=64, num_dims=3, do_code_regularization=True,
                          **util.SPECS["NetworkSpecs"], **util.SPECS["SubnetSpecs"])
            dec.load_state_dict(sd)
            dec.precision = "fp16x2"
            dec = dec.cuda().eval()
            pack = dec.pack(torch.device("cuda", 0))
            res = {}
            for four in (None, "1"):
                setenv(DEEPJOIN_B200_DEBUG="64", DEEPJOIN_B200_X2_FOURPASS=four)
                dec.query(code.cuda(), pts.cuda(), use_net=1)
                torch.cuda.synchronize()
                buf = np.zeros(64, np.uint64)
                _lib.check(_lib.lib().dj_debug_read_prof(pack.handle, buf.ctypes.data, buf.size))
                res[four] = buf.view(np.float64)[32 + 1]
            line.append("%8.1e" % (abs(res[None] - res["1"]) / max(abs(res["1"]), 1e-30)))
        print("ng %d: %s" % (ng, " ".join(line)), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "blocks":
    blocks()
