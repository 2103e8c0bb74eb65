"""three forward + backward passes of the attention head's training loop at the configured sizes (H = 512, 32 steps, N from argv,
default 32) and two greedy decodes (for ncu captures): python benchmarks/attn_train_once.py [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megreader_b200.refapi.decoders as md  # noqa: E402
from megreader_b200 import attn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = md.AttentionDecoder(256, inner_channels=512, max_size=32, height=1, gt_as_output=True).to(dev).train()
feat = torch.randn(n, 512, 1, 32, device=dev, requires_grad=True)
targets = torch.randint(1, 38, (n, 32), device=dev)
lengths = torch.randint(1, 30, (n,), device=dev)
fb = tuple(t.to(dev) for t in m.draw_feedback(n))
for _ in range(3):
    grid = torch.cat([feat, m._positions(n, dev)], dim=1)
    memory = grid.reshape(n, grid.shape[1], -1).permute(2, 0, 1)
    memory_bt = memory.transpose(0, 1)
    projected = m.decoder.attn.project_encoder(memory)
    loss, _ = attn.attention_loop_loss(projected, memory_bt, m.decoder, targets, lengths, fb, m.charset.blank)
    torch.autograd.grad(loss.mean(), [feat] + list(m.decoder.parameters()))
with torch.no_grad():
    m.eval()
    for _ in range(2):
        m._decode_cuda(memory_bt.detach(), projected.detach())
torch.cuda.synchronize()
print("ok")
