"""Latent step of the DeepSDF baseline (second net skipped) vs DeepJoin's, 8 shapes x 30,000 samples per launch."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR
from deepjoin_b200.networks.decoder_deepsdf import Decoder
SPECS = dict(dims=[512] * 8, dropout=list(range(8)), dropout_prob=0.2, norm_layers=list(range(8)), latent_in=[4],
             xyz_in_all=False, use_tanh=False, latent_dropout=False, weight_norm=True)
dev = torch.device("cuda", 0)
dj = make_decoder(0, dev, "bf16")
ds = Decoder(128, **SPECS); ds.load_state_dict(synth_data.make_deepsdf_state_dict(0, "trained_norm")); ds = ds.cuda().eval()
S, iters, ns = 8, 200, 30000
pools = []
for s in range(S):
    x, d, n = synth_data.make_sample_set(100000, seed=s)
    pools.append(tuple(torch.from_numpy(a.astype(np.float32)).to(dev) for a in (x, d.reshape(-1), n)))
sched = torch.randint(0, 100000, (S, iters, ns), device=dev, dtype=torch.int32)
for name, dec, C, kind, prec, flop in (("deepjoin bf16", dj, 192, 0, "bf16", 11.03e6), ("deepjoin fp16 forward", dj, 192, 0, "fp16", 11.03e6),
                                       ("deepsdf fp16 forward", ds, 129, 1, "fp16", 2 * 1706496 + 2 * 1703424)):
    codes0 = torch.randn(S, C) * 0.01
    if kind: codes0[:, 128] = 0
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        logs, codes = RR.optimize_codes(dec, codes0, pools, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=prec, loss_kind=kind)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-24s %.3f ms per shape-iteration, %.0f TFLOP/s algorithmic" % (name, ms / (S * iters), S * iters * ns * flop / (ms * 1e-3) / 1e12))
