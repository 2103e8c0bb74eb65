"""2D-CTC kernel micro-benchmark (cfg-3 shape T32 H8 C38 S32) — CUDA-event timings, L2 flushed between
iterations.  Prints one JSON line per (N, kernel, fast_math).  Run on the GPU box:
    python benchmarks/ctc2d_micro.py [N ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megreader_b200 import ctc2d  # noqa: E402
from tests.cases import ctc2d_case  # noqa: E402

T, H, C, S = 32, 8, 38, 32


def make(N, dev):
    base = min(N, 256)
    lp, tg, il, tl = ctc2d_case(3, T, H, base, C, S, 12)
    rep = (N + base - 1) // base
    lp = np.tile(lp, (1, 1, rep, 1))[:, :, :N]
    tg = np.tile(tg, (rep, 1))[:N]
    il = np.tile(il, rep)[:N]
    tl = np.tile(tl, rep)[:N]
    return [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (lp, tg, il, tl)]


def timeit(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def timeit_graph(fn, iters=20):
    """device time of one call without the host launch gap: `iters` calls captured in a CUDA graph, replayed once"""
    fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    dev = torch.device("cuda:0")
    Ns = [int(a) for a in sys.argv[1:]] or [32, 256, 2048, 16384]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lp_b, al_b, idx_b = T * H * C * 4, T * H * (2 * S + 1) * 4, 8 * S + 16
    for N in Ns:
        lp, tg, il, tl = make(N, dev)
        go = (1.0 / tl.float())
        for fast in (False, True):
            ctc2d.FAST_MATH = fast
            nll, la = ctc2d.ctc2d_forward(lp, tg, il, tl, 0, 0.0)
            _, gfac = ctc2d.ctc2d_forward_train(lp, tg, il, tl, 0)
            kernels = {
                "forward(alpha)": (lambda: ctc2d.ctc2d_forward(lp, tg, il, tl, 0, 0.0), lp_b + al_b + idx_b + 4),
                "backward(contract)": (lambda: ctc2d.ctc2d_backward(go, lp, tg, il, tl, nll, la, 0), 2 * lp_b + idx_b + 8),
                "forward_train": (lambda: ctc2d.ctc2d_forward_train(lp, tg, il, tl, 0), lp_b + T * C * 4 + idx_b + 4),
                "backward_apply": (lambda: ctc2d.ctc2d_backward_apply(go, lp, gfac), 2 * lp_b + T * C * 4 + 4),
            }
            for name, (fn, bytes_per_sample) in kernels.items():
                med, best = timeit(fn, flush)
                print(json.dumps({"N": N, "kernel": name, "fast_math": fast, "us_median": round(med, 2),
                                  "us_graph": round(timeit_graph(fn), 2) if N <= 2048 else None,
                                  "us_min": round(best, 2), "alg_bytes_per_sample": bytes_per_sample,
                                  "GBps_median": round(N * bytes_per_sample / med / 1e3, 1)}), flush=True)
        del lp, la, gfac


if __name__ == "__main__":
    main()
