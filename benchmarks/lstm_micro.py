"""BiLSTM recurrence timings (decoders/crnn.py BidirectionalLSTM), CUDA-event timed on one GPU.
    python benchmarks/lstm_micro.py [N I H]
  * kernel lines: the persistent whole-sequence kernels alone at T = 13 / 26 / 65 (slope = time per recurrent step)
  * layer lines: forward+backward of the whole layer (projection GEMMs included) per recurrence mode, graph-replayed."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from megreader_b200 import crnn_engine  # noqa: E402
from megreader_b200 import nnops as ops  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def trace(fn, T, dev, name):
    """median clock deltas between the stamps of CTA (0,0,0), see mr_lstm_seq_set_trace"""
    from megreader_b200 import _lib
    buf = torch.zeros(T, 32, dtype=torch.int64, device=dev)
    _lib.lib().mr_lstm_seq_set_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    _lib.lib().mr_lstm_seq_set_trace(None)
    t = buf.cpu().double()
    steps = slice(2, T)
    names = ["arrive->seen", "seen->tma_issued", "tma->mma_commit", "mma->acc_in_regs", "regs->stores", "stores->bar",
             "bar->fence", "fence->posted"]
    prev_post = t[1:T - 1, 7]
    d = {names[0]: (t[steps, 0] - prev_post)}
    for i in range(1, 8):
        d[names[i]] = t[steps, i] - t[steps, i - 1]
    d["period"] = t[steps, 7] - prev_post
    print(json.dumps({"bench": "lstm_seq_trace", "kernel": name, "unit": "SM clocks (median over steps)",
                      **{k: float(v.median()) for k, v in d.items()}}), flush=True)
    if name == "bwd":      # operand stream detail: k-block arrival (MMA warp) and issue (producer) times after "seen"
        arr = [float((t[steps, 8 + k] - t[steps, 0]).median()) for k in range(16)]
        iss = [float((t[steps, 24 + k] - t[steps, 0]).median()) for k in range(8)]
        print(json.dumps({"bench": "lstm_seq_trace_stream", "kblock_arrival_after_seen": arr,
                          "kblock_4_to_11_issue_after_seen": iss}), flush=True)


def kernels(N, H, dev):
    for T in (13, 26, 65):
        torch.manual_seed(T)
        Whh = [(torch.randn(4 * H, H, device=dev) / H ** 0.5).bfloat16() for _ in range(2)]
        bias = [torch.randn(4 * H, device=dev) * 0.1 for _ in range(2)]
        G0 = torch.randn(2, T, N, 4 * H, device=dev).bfloat16()
        G = G0.clone()
        C = torch.empty(2, T, N, H, device=dev)
        Y = torch.empty(T, N, 2 * H, device=dev, dtype=torch.bfloat16)
        dY = torch.randn(T, N, 2 * H, device=dev).bfloat16()
        dG = torch.empty_like(G)
        WhhT = [w.t().contiguous() for w in Whh]
        flags = ops.lstm_seq_flags(N, dev)
        ms_f = timed(lambda: ops.lstm_seq_fwd_tc(Whh, G, bias, C, Y, flags))
        ef = int(flags[-1])
        G.copy_(G0)
        ops.lstm_seq_fwd_tc(Whh, G, bias, C, Y, flags)
        ms_b = timed(lambda: ops.lstm_seq_bwd_tc(WhhT, G, C, dY, dG, flags))
        if T == 65:
            trace(lambda: ops.lstm_seq_fwd_tc(Whh, G, bias, C, Y, flags), T, dev, "fwd")
            trace(lambda: ops.lstm_seq_bwd_tc(WhhT, G, C, dY, dG, flags), T, dev, "bwd")
        print(json.dumps({"bench": "lstm_seq_kernels", "T": T, "N": N, "H": H, "fwd_ms": ms_f, "bwd_ms": ms_b,
                          "fwd_us_per_step": 1e3 * ms_f / T, "bwd_us_per_step": 1e3 * ms_b / T,
                          "err_word": [ef, int(flags[-1])]}), flush=True)


def layers(T, N, I, H, dev):
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.rnn = torch.nn.LSTM(I, H, bidirectional=True)
            self.embedding = torch.nn.Linear(2 * H, H)
    m = M().to(dev)
    x = torch.randn(T, N, I, device=dev)
    dout = torch.randn(T, N, H, device=dev)
    crnn_engine.set_compute_dtype(torch.bfloat16)
    for mode in ("cublas", "step", "seq"):
        crnn_engine.LSTM_MODE = mode

        def step():
            xe = x.clone().requires_grad_(True)
            out = crnn_engine.bilstm_forward(m, xe)
            out.float().backward(dout)
        for _ in range(3):
            step()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            step()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                step()
        torch.cuda.synchronize()
        ms = timed(g.replay)
        err = int(crnn_engine.LAST_LSTM_FLAGS[-1]) if crnn_engine.LAST_LSTM_FLAGS is not None else -1
        print(json.dumps({"bench": "bilstm_layer_fwd_bwd", "mode": mode, "T": T, "N": N, "I": I, "H": H, "ms": ms,
                          "graph": True, "err_word": err}), flush=True)


def main():
    N, I, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 512, 256)
    dev = torch.device("cuda:0")
    kernels(N, H, dev)
    layers(65, N, I, H, dev)


if __name__ == "__main__":
    main()
