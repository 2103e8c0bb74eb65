"""CPU: the phase-by-phase restatement of the attention head's training loop (oracle/attn_port.py = the arithmetic of
csrc/attn_decode.cu) against the framework composition of the reference-named modules (pinned to the unmodified reference by
tests/test_surfaces_cpu.py) and autograd of it, in float64."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("case", [(7, 32, 8, 2, 1, None, 0.2), (4, 48, 6, 1, 2, False, 0.0), (5, 32, 5, 1, 3, True, 0.4)],
                         ids=["coin-drop-2rows", "argmax", "teacher-drop"])
def test_restatement_matches_framework_loop_and_autograd(case):
    import megreader_b200.refapi.decoders as md
    from oracle import attn_port
    n, inner, S, height, seed, gt, drop = case
    torch.manual_seed(seed)
    m = md.AttentionDecoder(32, inner_channels=inner, max_size=S, height=height, gt_as_output=gt, step_dropout=drop).double().train()
    cell = m.decoder
    with torch.no_grad():
        for name, p in cell.named_parameters():
            if name.startswith(("rnn.", "out.", "attn.attn", "word_linear")) and p.dim() > 1:
                p.mul_(2.0)
        cell.embedding.weight.add_(0.05 * torch.randn_like(cell.embedding.weight))
    V, blank, H = len(m.charset), m.charset.blank, inner
    g = torch.Generator().manual_seed(500 + seed)
    feat = torch.randn(n, inner, height, S, generator=g, dtype=torch.float64).requires_grad_(True)
    lengths = torch.randint(1, S - 1, (n,), generator=g)
    targets = torch.randint(1, V, (n, S), generator=g)
    np.random.seed(seed)
    torch.manual_seed(seed)
    coin, swap, noise = m.draw_feedback(n)
    gout = torch.rand(n, generator=g, dtype=torch.float64) + 0.5
    grid = torch.cat([feat, m._positions(n, feat.device).double()], dim=1)
    memory = grid.reshape(n, grid.shape[1], -1).permute(2, 0, 1)
    memory_bt = memory.transpose(0, 1)
    projected = cell.attn.project_encoder(memory)
    det = lambda t: t.detach()  # noqa: E731
    wordtab = cell.word_linear(cell.embedding.weight)
    args = [det(t) for t in (projected, memory_bt.contiguous(), cell.attn.attn.weight[:, :H], cell.attn.v, wordtab, cell.rnn.weight_ih,
                             cell.rnn.bias_ih, cell.rnn.weight_hh, cell.rnn.bias_hh, cell.out.weight, cell.out.bias)]
    loss, maps, st = attn_port.forward(*args, targets, lengths, coin, swap, noise, blank)
    # the framework loop on the symbols the restatement fed back (each of them re-checked against the reference rule)
    hidden = memory_bt.new_zeros(n, H)
    loss_r, maps_r = 0, []
    for t in range(S):
        logp, hidden, w = cell(st["word"][t], hidden, memory, True, projected, memory_bt)
        loss_r = loss_r + m.loss_function(logp, targets[:, t]) * (t <= lengths).double()
        maps_r.append(w)
        if t + 1 < S:
            expect = targets[:, t] if bool(coin[t]) else logp.argmax(1)
            expect = torch.where(swap[t] == 1, noise[t], expect)
            assert torch.equal(st["word"][t + 1], expect)
    assert torch.equal(st["word"][0], torch.full((n,), blank))
    np.testing.assert_allclose(loss.numpy(), loss_r.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(maps.numpy(), torch.cat(maps_r, 1).detach().numpy(), rtol=1e-12, atol=1e-14)
    got = attn_port.backward(gout, args[0], args[1], args[2], args[3], args[5], args[7], args[9], targets, lengths, st)
    ref = torch.autograd.grad((loss_r * gout).sum(), [projected, memory_bt, cell.attn.v, cell.word_linear.weight, cell.attn.attn.weight,
                                                       cell.rnn.weight_ih, cell.rnn.bias_ih, cell.rnn.weight_hh, cell.rnn.bias_hh,
                                                       cell.out.weight, cell.out.bias])
    # the word table is word_linear(embedding.weight): its gradient reaches word_linear.weight as dwordtab^T . embedding.weight
    got["dword_linear"] = got["dwordtab"].t() @ cell.embedding.weight.detach()
    pairs = [("dP", ref[0]), ("dM", ref[1]), ("dv", ref[2]), ("dword_linear", ref[3]), ("dWa_h", ref[4][:, :H]), ("dW_ih", ref[5]),
             ("db_ih", ref[6]), ("dW_hh", ref[7]), ("db_hh", ref[8]), ("dW_out", ref[9]), ("db_out", ref[10])]
    for name, r in pairs:
        err = float((got[name] - r).abs().max())
        assert err <= 1e-11 * (1.0 + float(r.abs().max())), (name, err)
