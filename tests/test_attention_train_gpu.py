"""GPU parity of the attention head's TRAINING loop on the persistent kernels (csrc/attn_decode.cu, megreader_b200/attn.py) against the
framework composition of the same loop (decoders/attention_decoder.py:96-117, :187-231): per-sample loss, attention maps and the
gradients of every parameter of the cell and of the encoder grid, fp32.

The symbol a step feeds back may be the step's own argmax (a detached, discrete choice): a near-tie can fall either way under fp32
re-association and changes everything after it.  So the comparison arm replays the loop with the symbols the kernel fed back, and the
test separately checks that each of those symbols is what the reference rule selects from the comparison arm's own step outputs
(target under teacher forcing, dropout noise where drawn, otherwise an argmax up to 1e-5 in log-probability)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(in_ch, inner, max_size, height, seed, **kw):
    import megreader_b200.refapi.decoders as md
    torch.manual_seed(seed)
    m = md.AttentionDecoder(in_ch, inner_channels=inner, max_size=max_size, height=height, **kw)
    with torch.no_grad():
        for name, p in m.decoder.named_parameters():
            if name.startswith(("rnn.", "out.", "attn.attn", "word_linear")) and p.dim() > 1:
                p.mul_(2.0)
        m.decoder.embedding.weight.add_(0.05 * torch.randn_like(m.decoder.embedding.weight))
    return m.train()


def _loop_reference(m, memory, memory_bt, projected, words, targets, lengths):
    """the framework loop (AttentionRNNCell.forward per step) fed with a given symbol sequence words (S,N)"""
    n = memory_bt.shape[0]
    hidden = memory_bt.new_zeros(n, m.inner_channels)
    loss, maps, logps = 0, [], []
    for t in range(m.max_size):
        logp, hidden, w = m.decoder(words[t].long(), hidden, memory, True, projected, memory_bt)
        loss = loss + m.loss_function(logp, targets[:, t]) * (t <= lengths).float()
        maps.append(w)
        logps.append(logp)
    return loss, torch.cat(maps, 1), torch.stack(logps)


CASES = [  # n, inner, max_size, height, seed, gt_as_output, step_dropout
    (5, 128, 16, 2, 0, True, 0.0),
    (33, 512, 32, 1, 1, None, 0.2),       # the configured head: random teacher-forcing coin per step + step dropout
    (70, 256, 8, 1, 2, False, 0.0),       # always the step's own argmax
    (150, 512, 32, 1, 3, True, 0.3),
]


@pytest.mark.parametrize("case", CASES, ids=["n5-h128-2rows-tf", "n33-h512-coin-drop", "n70-h256-argmax", "n150-h512-tf-drop"])
def test_training_loop_matches_framework_loop(cuda, case):
    from megreader_b200 import attn as attn_kernels
    n, inner, max_size, height, seed, gt, drop = case
    m = _model(32, inner, max_size, height, seed, gt_as_output=gt, step_dropout=drop).to(cuda)
    vocab, blank = len(m.charset), m.charset.blank
    g = torch.Generator().manual_seed(500 + seed)
    feat = torch.randn(n, inner, height, max_size, generator=g).to(cuda).requires_grad_(True)
    lengths = torch.randint(1, max_size - 1, (n,), generator=g)
    targets = torch.randint(1, vocab, (n, max_size), generator=g)
    for b in range(n):
        targets[b, lengths[b]:] = blank
    targets, lengths = targets.to(cuda), lengths.to(cuda)
    np.random.seed(seed)
    torch.manual_seed(seed)
    feedback = tuple(t.to(cuda) for t in m.draw_feedback(n))
    coin, swap, noise = feedback
    gout = torch.rand(n, generator=g).to(cuda) + 0.5

    def front():
        grid = torch.cat([feat, m._positions(n, feat.device)], dim=1)
        memory = grid.reshape(n, grid.shape[1], -1).permute(2, 0, 1)
        return memory, memory.transpose(0, 1), m.decoder.attn.project_encoder(memory)

    params = [p for p in m.decoder.parameters()]
    names = [k for k, _ in m.decoder.named_parameters()]
    # kernels
    memory, memory_bt, projected = front()
    loss_k, maps_k, words = attn_kernels.attention_loop_loss(projected, memory_bt, m.decoder, targets, lengths, feedback, blank,
                                                             return_words=True)
    grads_k = torch.autograd.grad((loss_k * gout).sum(), [feat] + params, allow_unused=True)
    # framework loop on the same symbols
    memory, memory_bt, projected = front()
    loss_r, maps_r, logp_r = _loop_reference(m, memory, memory_bt, projected, words, targets, lengths)
    grads_r = torch.autograd.grad((loss_r * gout).sum(), [feat] + params, allow_unused=True)

    # the fed-back symbols follow the reference rule (attention_decoder.py:106-116)
    assert tuple(words.shape) == (max_size, n) and bool((words[0] == blank).all())
    w_next = words[1:].long()
    lp = logp_r[:-1].detach()
    chosen = lp.gather(2, w_next.unsqueeze(2)).squeeze(2)
    is_argmax = chosen >= lp.max(dim=2).values - 1e-5
    is_target = w_next == targets[:, :max_size - 1].t()
    is_noise = (swap[:-1] == 1) & (w_next == noise[:-1])
    teacher = coin[:-1].view(-1, 1).expand_as(w_next)
    ok = torch.where(swap[:-1] == 1, is_noise, torch.where(teacher, is_target, is_argmax))
    assert bool(ok.all()), "%d fed-back symbols break the reference rule" % int((~ok).sum())
    if gt is False:
        assert int((w_next != targets[:, :max_size - 1].t()).sum()) > 0          # the argmax path was really taken

    np.testing.assert_allclose(loss_k.detach().cpu().numpy(), loss_r.detach().cpu().numpy(), rtol=2e-5, atol=1e-4)
    np.testing.assert_allclose(maps_k.cpu().numpy(), maps_r.detach().cpu().numpy(), rtol=1e-4, atol=1e-6)
    for name, gk, gr in zip(["feature"] + names, grads_k, grads_r):
        assert (gk is None) == (gr is None), name
        if gk is None:
            continue
        scale = float(gr.abs().max())
        err = float((gk - gr).abs().max())
        assert err <= 2e-4 * scale + 1e-7, "%s: max abs error %.3e at scale %.3e" % (name, err, scale)


def test_module_training_forward_uses_the_kernels(cuda):
    """AttentionDecoder.forward in training mode on CUDA = the kernels (same draws as the framework composition of the module)"""
    n, inner, max_size, height = 12, 128, 16, 1
    m = _model(32, inner, max_size, height, 7, gt_as_output=True, step_dropout=0.0).to(cuda)
    x = torch.randn(n, 32, 16, 2 * max_size, device=cuda)
    lengths = torch.randint(1, max_size - 1, (n,), device=cuda)
    targets = torch.randint(1, len(m.charset), (n, max_size), device=cuda)
    from megreader_b200 import _lib
    out = {}
    for mode in (True, False):
        m.loop_kernels = mode
        m.zero_grad()
        torch.manual_seed(3)
        _lib.reset_launch_count()
        loss, maps = m(x, targets=targets, lengths=lengths, train=True)
        loss.mean().backward()
        out[mode] = (loss.detach().clone(), maps.detach().clone(), m.decoder.rnn.weight_ih.grad.clone(),
                     m.encode[0][0].weight.grad.clone(), _lib.launch_count())
    assert out[True][4] >= 2 and out[False][4] == 0          # forward + backward kernel vs no repo kernel at all
    assert tuple(out[True][1].shape) == (n, max_size, height, max_size)
    for a, b in zip(out[True][:4], out[False][:4]):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 5e-4 * scale + 1e-7
