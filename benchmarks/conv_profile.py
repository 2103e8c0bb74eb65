"""Launch the dominant CRNN kernels in isolation at the batch-512 shapes (for ncu --set full and for timing).
    python benchmarks/conv_profile.py            # conv5 fprop / dgrad / wgrad x3"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megreader_b200 import nnops  # noqa: E402

dev = torch.device("cuda:0")
N, H, W, C, Cout, k, p = 512, 4, 65, 512, 512, 3, 1       # conv5 of backbones/crnn.py at 32x256 input, batch 512
torch.manual_seed(0)
x = torch.randn(N, H, W, C, device=dev).bfloat16()
wm = (torch.randn(Cout, k * k * C, device=dev) / 68).bfloat16()
dz = torch.randn(N, H, W, Cout, device=dev).bfloat16()
for _ in range(3):
    y, Ho, Wo = nnops.conv_fprop_tc(x, wm, k, k, p, p)
    dx, _, _ = nnops.conv_fprop_tc(dz, wm, k, k, k - 1 - p, k - 1 - p)
    dw = nnops.conv_wgrad_tc(dz, x, k, k, p, p)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    y, Ho, Wo = nnops.conv_fprop_tc(x, wm, k, k, p, p)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) * 100
print("conv5 fprop %.1f us  %.1f TFLOP/s" % (us, 2.0 * N * H * W * Cout * k * k * C / us / 1e6))
# conv1 (low-channel, epilogue-heavy): 512 x 16 x 128 x 64 -> 128
x1 = torch.randn(512, 16, 128, 64, device=dev).bfloat16()
w1 = (torch.randn(128, 9 * 64, device=dev) / 24).bfloat16()
nnops.conv_fprop_tc(x1, w1, 3, 3, 1, 1)
torch.cuda.synchronize()
a.record()
for _ in range(10):
    nnops.conv_fprop_tc(x1, w1, 3, 3, 1, 1)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) * 100
print("conv1 fprop %.1f us  %.1f TFLOP/s" % (us, 2.0 * 512 * 16 * 128 * 128 * 9 * 64 / us / 1e6))
