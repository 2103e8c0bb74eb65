"""GPU parity of the persistent greedy-decoding kernel of the attention head (csrc/attn_decode.cu) against the framework composition
of the same loop (decoders/attention_decoder.py:119-131, :187-231): identical symbols, per-step softmax to 1e-4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(in_ch, inner, max_size, height, seed):
    import megreader_b200.refapi.decoders as md
    torch.manual_seed(seed)
    m = md.AttentionDecoder(in_ch, inner_channels=inner, max_size=max_size, height=height)
    with torch.no_grad():                           # livelier recurrences than the default init: logits that actually move
        for name, p in m.decoder.named_parameters():
            if name.startswith(("rnn.", "out.", "attn.attn", "word_linear")) and p.dim() > 1:
                p.mul_(3.0)
        m.decoder.embedding.weight.add_(0.05 * torch.randn_like(m.decoder.embedding.weight))   # not the identity any more
    return m.eval()


@pytest.mark.parametrize("case", [(3, 64, 128, 16, 2, 0), (70, 32, 512, 32, 1, 1), (200, 16, 256, 8, 1, 2)],
                         ids=["n3-h128-2rows", "n70-h512", "n200-h256"])
def test_decode_kernel_matches_framework_loop(cuda, case):
    n, in_ch, inner, max_size, height, seed = case
    m = _model(in_ch, inner, max_size, height, seed).to(cuda)
    hh = {1: 16, 2: 32}[height]                       # encoder: pools (2,2),(2,1),(2,1) then the (2,3)/(2,1) conv -> height rows
    x = torch.randn(n, in_ch, hh, 2 * max_size, device=cuda)
    with torch.no_grad():
        grid = torch.cat([m.encode(x).float(), m._positions(n, x.device)], dim=1)
        assert grid.shape[2] == height and grid.shape[3] == max_size
        memory = grid.reshape(n, grid.shape[1], -1).permute(2, 0, 1)
        memory_bt = memory.transpose(0, 1)
        projected = m.decoder.attn.project_encoder(memory)
        pred_ref, prob_ref = m._decode_aten(memory, memory_bt, projected, want_prob=True)
        pred, prob = m._decode_cuda(memory_bt, projected, want_prob=True)
    assert pred.dtype == torch.int32 and tuple(pred.shape) == (n, max_size)
    pr, pk = pred_ref.cpu().numpy(), pred.cpu().numpy()
    assert len(np.unique(pr)) > 3, "degenerate case: the loop emits (almost) one symbol"
    assert np.array_equal(pr, pk), "symbols differ at %d of %d positions" % ((pr != pk).sum(), pr.size)
    np.testing.assert_allclose(prob.cpu().numpy(), prob_ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    # the module's eval forward = the kernel + the reference's early exit
    with torch.no_grad():
        out = m(x)
    blank = m.charset.blank
    fin = np.maximum.accumulate((pr == blank).all(axis=0))
    expect = np.where(fin[None, :], blank, pr)
    assert np.array_equal(out.cpu().numpy(), expect)
