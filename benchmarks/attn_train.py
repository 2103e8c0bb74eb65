"""Training loop of the attention head (fwd + bwd through time): persistent kernels (csrc/attn_decode.cu) vs the framework composition
of the same 32 steps, both replayed from a CUDA graph (device time without host launch gaps) and eager."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megreader_b200.refapi.decoders as md  # noqa: E402


def timed(fn, iters=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = md.AttentionDecoder(256, inner_channels=512, max_size=32, height=1, gt_as_output=True).to(dev).train()
    for n in (32, 128, 256):
        feat = torch.randn(n, 512, 1, 32, device=dev, requires_grad=True)
        targets = torch.randint(1, 38, (n, 32), device=dev)
        lengths = torch.randint(1, 30, (n,), device=dev)
        m.feedback_static = tuple(t.to(dev) for t in m.draw_feedback(n))
        params = list(m.decoder.parameters())

        def step():
            grid = torch.cat([feat, m._positions(n, dev)], dim=1)
            memory = grid.reshape(n, grid.shape[1], -1).permute(2, 0, 1)
            memory_bt = memory.transpose(0, 1)
            projected = m.decoder.attn.project_encoder(memory)
            if m.loop_kernels:
                from megreader_b200 import attn
                loss, _ = attn.attention_loop_loss(projected, memory_bt, m.decoder, targets, lengths, m.feedback_static, m.charset.blank)
            else:
                hidden = memory_bt.new_zeros(n, 512)
                word = torch.full((n,), m.charset.blank, dtype=torch.long, device=dev)
                loss = 0
                for t in range(32):
                    logp, hidden, _ = m.decoder(word, hidden, memory, True, projected, memory_bt)
                    loss = loss + m.loss_function(logp, targets[:, t]) * (t <= lengths).float()
                    word = targets[:, t]
            return torch.autograd.grad(loss.mean(), [feat] + params)
        res = {}
        for mode in (True, False):
            m.loop_kernels = mode
            res[mode] = (timed(step), timed(graphed(step)))
        print("N=%d: kernels %.3f ms eager / %.3f ms graph | framework loop %.3f ms eager / %.3f ms graph | graph speed-up %.2fx"
              % (n, res[True][0], res[True][1], res[False][0], res[False][1], res[False][1] / res[True][1]), flush=True)


if __name__ == "__main__":
    main()
