"""greedy decoding loop of the attention head: persistent kernel vs the framework composition: python benchmarks/attn_decode.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megreader_b200.refapi.decoders as md
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = md.AttentionDecoder(256).to(dev).eval()
for n in (16, 64, 256, 1024):
    mem_bt = torch.randn(n, 32, 512 + 33, device=dev)
    memory = mem_bt.transpose(0, 1)
    with torch.no_grad():
        projected = m.decoder.attn.project_encoder(memory)
        res = {}
        for name, fn in (("kernel", lambda: m._decode_cuda(mem_bt, projected)), ("aten", lambda: m._decode_aten(memory, mem_bt, projected))):
            for _ in range(2):
                out = fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                out = fn()
            b.record()
            torch.cuda.synchronize()
            res[name] = (a.elapsed_time(b) / 5, out)
    same = bool((res["kernel"][1].long() == res["aten"][1].long()).all())
    print("N=%d: kernel %.3f ms (%.1f k lines/s) | framework loop %.3f ms | %.1fx | same symbols: %s" % (
        n, res["kernel"][0], n / res["kernel"][0], res["aten"][0], res["aten"][0] / res["kernel"][0], same), flush=True)
