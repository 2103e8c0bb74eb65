"""DCNv2 backward at the bench shapes, fused tcgen05 weight gradient vs the im2col + SGEMM route: python benchmarks/dcn_bwd.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megreader_b200 import dcn
dev = torch.device("cuda:0")
B = 8
for C, H in ((128, 64), (256, 32), (512, 16)):
    torch.manual_seed(0)
    x = torch.randn(B, C, H, H, device=dev)
    w = torch.randn(C, C, 3, 3, device=dev) / (3 * C ** 0.5)
    off = 2 * torch.randn(B, 18, H, H, device=dev)
    m = torch.sigmoid(torch.randn(B, 9, H, H, device=dev))
    go = torch.randn(B, C, H, H, device=dev)
    res = {}
    for mode in ("fused", "unfused"):
        os.environ.pop("MR_DCN_UNFUSED_WGRAD", None)
        if mode == "unfused":
            os.environ["MR_DCN_UNFUSED_WGRAD"] = "1"
        gi, gw, goff, gm = torch.zeros_like(x), torch.zeros_like(w), torch.zeros_like(off), torch.zeros_like(m)
        def run(only_w=False):
            dcn.modulated_deform_conv_cuda_backward(x, w, None, None, off, m, None, None if only_w else gi, gw, None,
                                                    None if only_w else goff, None if only_w else gm, go, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, False)
        for only_w in (False, True):
            for _ in range(3):
                run(only_w)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                run(only_w)
            b.record()
            torch.cuda.synchronize()
            res[(mode, only_w)] = a.elapsed_time(b) / 10 * 1e3
        gw.zero_(); run(True); res[mode + "_gw"] = gw.clone()
    err = float((res["fused_gw"] - res["unfused_gw"]).abs().max() / res["unfused_gw"].abs().max())
    print("C%d@%d: bwd fused %.1f us unfused %.1f us | wgrad only fused %.1f unfused %.1f | gw rel diff %.2e" % (
        C, H, res[("fused", False)], res[("unfused", False)], res[("fused", True)], res[("unfused", True)], err), flush=True)
