"""BiLSTM layer (decoders/crnn.py BidirectionalLSTM) forward+backward time per recurrence mode, CUDA-event timed.
    python benchmarks/lstm_micro.py [T N I H]   ->  one JSON line per mode"""
import json
import sys

import torch

from megreader_b200 import crnn_engine


def main():
    T, N, I, H = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (26, 512, 512, 256)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)

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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            g.replay()
        e0.record()
        reps = 20
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        err = int(crnn_engine.LAST_LSTM_FLAGS[-1]) if crnn_engine.LAST_LSTM_FLAGS is not None else -1
        print(json.dumps({"bench": "bilstm_layer_fwd_bwd", "mode": mode, "T": T, "N": N, "I": I, "H": H,
                          "ms": e0.elapsed_time(e1) / reps, "graph": True, "err_word": err}), flush=True)


if __name__ == "__main__":
    main()
