"""DCNv2 micro-benchmark (SURVEY.md §8d shapes: (C,H) = (128,64), (256,32), (512,16), B = 8, 3x3, dg = 1, fp32):
forward and backward through megreader_b200.dcn, CUDA-event timed, with the algorithmic column bytes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megreader_b200 import dcn  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for C, H in ((128, 64), (256, 32), (512, 16)):
    torch.manual_seed(0)
    x = torch.randn(B, C, H, H, device=dev, requires_grad=True)
    w = (torch.randn(C, C, 3, 3, device=dev) / (3 * C ** 0.5)).requires_grad_(True)
    off = (2 * torch.randn(B, 18, H, H, device=dev)).requires_grad_(True)
    m = torch.sigmoid(torch.randn(B, 9, H, H, device=dev)).requires_grad_(True)
    go = torch.randn(B, C, H, H, device=dev)

    def fwd():
        return dcn.modulated_deform_conv(x, off, m, w, None, 1, 1, 1, 1, 1)

    def fwdbwd():
        for t in (x, w, off, m):
            t.grad = None
        fwd().backward(go)
    res = {"B": B, "C": C, "H": H}
    for name, fn in (("fwd", fwd), ("fwd+bwd", fwdbwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        res[name + "_us"] = a.elapsed_time(b) * 100
    col_bytes = 9 * C * H * H * 4 * B
    res["col_MB"] = col_bytes / 1e6
    res["gemm_GFLOP_fwd"] = 2.0 * B * C * 9 * C * H * H / 1e9
    print(json.dumps(res), flush=True)
