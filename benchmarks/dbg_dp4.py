"""debug: v4 DP kernel vs v3 on the full-size property case; prints the mismatching (t, n)."""
import os, sys, subprocess, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megreader_b200 import ctc2d
from tests.cases import ctc2d_case
T, H, N, C, S = 32, 8, 4096, 38, 32
lp, tg, il, tl = ctc2d_case(3, T, H, N, C, S, 12)
dev = torch.device("cuda:0")
d = [torch.from_numpy(a).to(dev) for a in (lp, tg, il, tl)]
mode = sys.argv[1]
nll, fac = ctc2d.ctc2d_forward_train(*d, 0)
np.savez("gpurun_out/dp4_%s.npz" % mode, nll=nll.cpu().numpy(), fac=fac.cpu().numpy())
if mode == "v4":
    a = np.load("gpurun_out/dp4_v3.npz")
    f3, f4 = a["fac"], fac.cpu().numpy()
    print("nll maxrel", np.abs(a["nll"] - nll.cpu().numpy()).max() / np.abs(a["nll"]).max())
    bad = np.argwhere(np.abs(f3 - f4) > 1e-4)
    print("mismatches", len(bad))
    for t, n, c in bad[:20]:
        print("t", t, "n", n, "c", c, "v3", f3[t, n, c], "v4", f4[t, n, c], "L", tl[n], "targets", tg[n, :tl[n]].tolist())
