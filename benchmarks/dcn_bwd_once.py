"""two DCNv2 backward calls per bench shape (for ncu captures): python benchmarks/dcn_bwd_once.py"""
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
    gi, gw, goff, gm = torch.zeros_like(x), torch.zeros_like(w), torch.zeros_like(off), torch.zeros_like(m)
    for _ in range(2):
        dcn.modulated_deform_conv_cuda_backward(x, w, None, None, off, m, None, gi, gw, None, goff, gm, go, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, False)
    torch.cuda.synchronize()
