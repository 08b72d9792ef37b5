"""bf16 vs fp32 latent gradient of the DeepSDF objective at several batch sizes."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from deepjoin_b200 import reconstruct as R
from deepjoin_b200.networks.decoder_deepsdf import Decoder
SPECS = dict(dims=[512] * 8, dropout=list(range(8)), dropout_prob=0.2, norm_layers=list(range(8)), latent_in=[4],
             xyz_in_all=False, use_tanh=False, latent_dropout=False, weight_norm=True)
dec = Decoder(128, **SPECS); dec.load_state_dict(synth_data.make_deepsdf_state_dict(0, "trained_norm")); dec = dec.cuda().eval()
xyz, sdf, _ = synth_data.make_sample_set(60000, seed=2)
torch.manual_seed(1)
lat = torch.ones(1, 128).normal_(mean=0, std=0.01).reshape(-1)
for n in (128, 256, 512, 1024, 2000, 30000):
    x = torch.from_numpy(xyz[:n].astype(np.float32)).cuda(); s = torch.from_numpy(sdf[:n].reshape(-1).astype(np.float32)).cuda()
    out = {}
    for prec in ("fp32", "bf16", "fp16"):
        g, t = R.latent_grad(dec, dec.code192(lat), x, s, x, clamp_dist=0.1, l2reg=True, precision=prec, loss_kind=1)
        out[prec] = (g.cpu().numpy()[:128], t.cpu().numpy())
    a = out["fp32"][0]
    for prec in ("bf16", "fp16"):
        b = out[prec][0]
        print(n, prec, "cos %.4f |g32| %.4e |g| %.4e loss32 %.5f loss %.5f" % (np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)), np.linalg.norm(a), np.linalg.norm(b), out["fp32"][1][0], out[prec][1][0]))

# the DeepJoin objective on the two-net decoder
from bench import make_decoder
dj = make_decoder(0, torch.device("cuda", 0), "bf16")
xyz, sdf, nrm = synth_data.make_sample_set(60000, seed=2)
torch.manual_seed(1)
code = R.init_code(128, 64, 0.01)
for n in (512, 2000, 30000):
    f = lambda a: torch.from_numpy(a[:n].astype(np.float32)).cuda()
    out = {}
    for prec in ("fp32", "bf16", "fp16"):
        g, t = R.latent_grad(dj, code, f(xyz), f(sdf.reshape(-1)), f(nrm), precision=prec)
        out[prec] = (g.cpu().numpy(), t.cpu().numpy())
    a = out["fp32"][0]
    for prec in ("bf16", "fp16"):
        b = out[prec][0]
        print("deepjoin", n, prec, "cos %.4f relerr %.4f loss32 %.5f loss %.5f" % (np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)), np.linalg.norm(a - b) / np.linalg.norm(a), out["fp32"][1][0], out[prec][1][0]))
