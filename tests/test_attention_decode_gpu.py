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


def _first_undecided(prob_ref, thr):
    """per sample: first step whose top-2 softmax margin in the framework loop is below `thr` (max_size when none is): a near-tie may
    legitimately fall the other way under fp32 re-association, and everything after it is a different (equally valid) trajectory"""
    top2 = prob_ref.topk(2, dim=2).values
    close = ((top2[..., 0] - top2[..., 1]) < thr).cpu().numpy()
    return np.where(close.any(axis=1), close.argmax(axis=1), close.shape[1])


@pytest.mark.parametrize("case", [(3, 64, 128, 16, 2, 0), (70, 32, 512, 32, 1, 1), (200, 16, 256, 8, 1, 2)],
                         ids=["n3-h128-2rows", "n70-h512", "n200-h256"])
def test_decode_kernel_matches_framework_loop(cuda, case):
    """random encoder features (unit scale) so that every sample follows its own trajectory: symbols identical up to each sample's
    first near-tie, per-step softmax to 1e-4 on those steps"""
    n, in_ch, inner, max_size, height, seed = case
    m = _model(in_ch, inner, max_size, height, seed).to(cuda)
    g = torch.Generator().manual_seed(1000 + seed)
    feat = torch.randn(n, inner, height, max_size, generator=g).to(cuda)
    with torch.no_grad():
        grid = torch.cat([feat, m._positions(n, feat.device)], dim=1)
        memory = grid.reshape(n, grid.shape[1], -1).permute(2, 0, 1)
        memory_bt = memory.transpose(0, 1)
        projected = m.decoder.attn.project_encoder(memory)
        pred_ref, prob_ref = m._decode_aten(memory, memory_bt, projected, want_prob=True)
        pred, prob = m._decode_cuda(memory_bt, projected, want_prob=True)
    assert pred.dtype == torch.int32 and tuple(pred.shape) == (n, max_size)
    pr, pk = pred_ref.cpu().numpy(), pred.cpu().numpy()
    assert len(np.unique(pr)) > 3 and len(np.unique(pr, axis=0)) > min(n, 3) - 1, "degenerate case: the loop emits (almost) one string"
    stop = _first_undecided(prob_ref, 1e-4)
    valid = np.arange(max_size)[None, :] < stop[:, None]            # steps strictly before the first near-tie: symbols must agree
    assert valid.mean() >= 0.6, "too few positions before a near-tie: %s" % (stop,)
    assert np.array_equal(pr[valid], pk[valid]), "symbols differ at %d of %d decided positions" % ((pr != pk)[valid].sum(), valid.sum())
    upto = np.arange(max_size)[None, :] <= stop[:, None]            # the near-tie step itself still has the same history
    np.testing.assert_allclose(prob.cpu().numpy()[upto], prob_ref.cpu().numpy()[upto], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", [(3, 64, 128, 16, 2, 0), (20, 32, 512, 32, 1, 1)], ids=["n3-h128-2rows", "n20-h512"])
def test_eval_forward_is_kernel_plus_early_exit(cuda, case):
    """the module's eval forward on images (conv encoder included) = the kernel + the reference's early exit (attention_decoder.py:129-130)"""
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
        out = m(x)
    pr = pred_ref.cpu().numpy()
    stop = _first_undecided(prob_ref, 1e-3)
    if (stop < max_size).any():
        pytest.skip("a near-tie in the framework loop: trajectories may legitimately differ")
    blank = m.charset.blank
    fin = np.maximum.accumulate((pr == blank).all(axis=0))
    expect = np.where(fin[None, :], blank, pr)
    assert np.array_equal(out.cpu().numpy(), expect)
