"""one DCNv2 forward per shape (for ncu captures): python benchmarks/dcn_fwd_only.py [B]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megreader_b200 import dcn
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for C, H in ((128, 64), (256, 32), (512, 16)):
    torch.manual_seed(0)
    x = torch.randn(B, C, H, H, device=dev)
    w = torch.randn(C, C, 3, 3, device=dev) / (3 * C ** 0.5)
    off = 2 * torch.randn(B, 18, H, H, device=dev)
    m = torch.sigmoid(torch.randn(B, 9, H, H, device=dev))
    for _ in range(3):
        dcn.modulated_deform_conv(x, off, m, w, None, 1, 1, 1, 1, 1)
    torch.cuda.synchronize()
