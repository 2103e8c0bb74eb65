"""2D-CTC head epilogue micro-benchmark (cfg-3 shape C38 H8 W32): CUDA-event timings vs the ATen composition of
decoders/ctc_decoder2d.py:37-45.  One JSON line per (N, variant).   python benchmarks/ctc2d_head_micro.py [N ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megreader_b200 import ctc2d_head  # noqa: E402

C, H, W = 38, 8, 32


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    for N in ([int(a) for a in sys.argv[1:]] or [2048, 16384]):
        m = torch.randn(N, 1, H, W, device=dev)
        z = torch.randn(N, C, H, W, device=dev)
        dlp = torch.randn(W, H, N, C, device=dev)
        gfac = torch.randn(W, N, C, device=dev)
        go = torch.rand(N, device=dev)
        tiny = torch.tensor(torch.finfo(torch.float32).tiny, device=dev)
        lp_bytes = N * C * H * W * 4

        def aten_fwd():
            p = torch.softmax(m, 2) * torch.softmax(z, 1)
            return torch.log(torch.max(p, tiny)).permute(3, 2, 0, 1).contiguous()
        rows = [("fwd", timed(lambda: ctc2d_head.head_forward(m, z)), 2 * lp_bytes + m.numel() * 4),
                ("bwd_explicit", timed(lambda: ctc2d_head.head_backward(m, z, grad_lp=dlp)), 3 * lp_bytes),
                ("bwd_factored", timed(lambda: ctc2d_head.head_backward(m, z, gfac=gfac, grad_out=go)),
                 2 * lp_bytes + gfac.numel() * 4),
                ("aten_fwd", timed(aten_fwd), 2 * lp_bytes)]
        for name, us, nbytes in rows:
            print(json.dumps({"bench": "ctc2d_head", "N": N, "variant": name, "us": us, "algorithmic_bytes": nbytes,
                              "GBps": nbytes / us / 1e3}), flush=True)


if __name__ == "__main__":
    main()
