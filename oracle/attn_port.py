"""TEST INFRASTRUCTURE -- never imported by the product (megreader_b200/).  Phase-by-phase CPU restatement (plain PyTorch tensor
arithmetic, any dtype; the tests use float64) of what csrc/attn_decode.cu computes for the attention head's TRAINING loop, i.e. of
decoders/attention_decoder.py:96-117 (loop, teacher forcing :106-110, step dropout :111-116, masked NLL :104) around
AttentionRNNCell.forward (:187-231: embedding + word_linear, Attn.forward :146-171 = softmax_l(v . tanh(W [h ; enc_l] + b)), bmm context,
GRUCell, out Linear, log_softmax) and of its derivative written out by hand the way attn_bwd_kernel evaluates it (B1 .. B4).

Pin: tests/test_oracle_attn.py checks forward() against the framework composition of the reference-named modules (which
tests/test_surfaces_cpu.py pins to the unmodified reference) and backward() against autograd of that composition, to 1e-12 in float64.
"""
import torch


def forward(P, M, Wa_h, v, wordtab, W_ih, b_ih, W_hh, b_hh, W_out, b_out, targets, lengths, coin, swap, noise, blank):
    """P (N,L,H) = Wa_enc . memory + b, M (N,L,D) memory, targets (N,S), lengths (N), coin (S) bool, swap / noise (S,N).
    -> loss (N,), attn (N,S,L), saved state dict (what attn_fwd_kernel<true> keeps for the backward)."""
    N, L, H = P.shape
    S, V, X = coin.shape[0], W_out.shape[0], W_ih.shape[1]
    z = lambda *s: torch.zeros(*s, dtype=P.dtype)  # noqa: E731
    h_all, fh_all, x_all, gates = z(S + 1, N, H), z(S, N, H), z(S, N, X), z(S, N, 4, H)
    logp, attn, loss = z(S, N, V), z(N, S, L), z(N)
    word = torch.zeros(S, N, dtype=torch.long)
    for t in range(S + 1):
        h = h_all[t]
        if t < S:                                                    # P1
            fh_all[t] = h @ Wa_h.t()
        if t > 0:                                                    # P2, head of the previous step
            logits = h @ W_out.t() + b_out
            best = logits.max(1, keepdim=True).values
            logp[t - 1] = (logits - best) - (logits - best).exp().sum(1, keepdim=True).log()
            tgt = targets[:, t - 1]
            loss -= torch.where(t - 1 <= lengths, logp[t - 1].gather(1, tgt[:, None])[:, 0], z(N))
            w = tgt if bool(coin[t - 1]) else logits.argmax(1)
            w = torch.where(swap[t - 1] == 1, noise[t - 1], w)
        else:
            w = torch.full((N,), blank, dtype=torch.long)
        if t == S:
            break
        word[t] = w
        a = torch.softmax(torch.tanh(P + fh_all[t][:, None, :]) @ v, 1)   # P2, attention
        attn[:, t] = a
        x_all[t] = torch.cat([wordtab[w], (a[:, :, None] * M).sum(1)], 1)
        gi, gh = x_all[t] @ W_ih.t(), h @ W_hh.t()                    # P3, GRU cell (gate order r, z, n)
        r = torch.sigmoid(gi[:, :H] + b_ih[:H] + gh[:, :H] + b_hh[:H])
        zz = torch.sigmoid(gi[:, H:2 * H] + b_ih[H:2 * H] + gh[:, H:2 * H] + b_hh[H:2 * H])
        ghn = gh[:, 2 * H:] + b_hh[2 * H:]
        nn = torch.tanh(gi[:, 2 * H:] + b_ih[2 * H:] + r * ghn)
        h_all[t + 1] = (1 - zz) * nn + zz * h
        gates[t, :, 0], gates[t, :, 1], gates[t, :, 2], gates[t, :, 3] = r, zz, nn, ghn
    return loss, attn, dict(h_all=h_all, fh_all=fh_all, x_all=x_all, gates=gates, logp=logp, attn=attn, word=word)


def backward(gloss, P, M, Wa_h, v, W_ih, W_hh, W_out, targets, lengths, st):
    """gradients for the upstream gradient gloss (N,) of the loss: dict with dP, dM, dv, dwordtab, dWa_h, dW_ih, db_ih, dW_hh, db_hh,
    dW_out, db_out -- evaluated step by step in reverse like attn_bwd_kernel (B1 output layer + GRU cell, B2 gate products
    transposed, B3 attention, B4 hidden half of the energy), weight gradients as products over the S*N saved rows."""
    N, L, H = P.shape
    D, V = M.shape[2], W_out.shape[0]
    h_all, fh_all, x_all, gates, logp, attn, word = (st[k] for k in ("h_all", "fh_all", "x_all", "gates", "logp", "attn", "word"))
    S, X = fh_all.shape[0], x_all.shape[2]
    z = lambda *s: torch.zeros(*s, dtype=P.dtype)  # noqa: E731
    dh, dP, dM, dv, dwt = z(N, H), z(N, L, H), z(N, L, D), z(H), z(V, H)
    dlogits, dgi, dgh, dfh = z(S, N, V), z(S, N, 3 * H), z(S, N, 3 * H), z(S, N, H)
    for t in range(S - 1, -1, -1):
        msk = torch.where(t <= lengths, gloss, z(N))                                            # B1
        dlogits[t] = (logp[t].exp() - torch.nn.functional.one_hot(targets[:, t], V)) * msk[:, None]
        d = dh + dlogits[t] @ W_out
        r, zz, nn, ghn = gates[t, :, 0], gates[t, :, 1], gates[t, :, 2], gates[t, :, 3]
        dn_pre = d * (1 - zz) * (1 - nn * nn)
        dz_pre = d * (h_all[t] - nn) * zz * (1 - zz)
        dr_pre = dn_pre * ghn * r * (1 - r)
        dgi[t] = torch.cat([dr_pre, dz_pre, dn_pre], 1)
        dgh[t] = torch.cat([dr_pre, dz_pre, dn_pre * r], 1)
        dh = d * zz
        dx = dgi[t] @ W_ih                                                                       # B2
        dh = dh + dgh[t] @ W_hh
        a, dctx = attn[:, t], dx[:, H:]                                                          # B3
        dwt.index_add_(0, word[t], dx[:, :H])
        da = (dctx[:, None, :] * M).sum(2)
        ds = a * (da - (a * da).sum(1, keepdim=True))
        dM += a[:, :, None] * dctx[:, None, :]
        e = torch.tanh(P + fh_all[t][:, None, :])
        dv += (ds[:, :, None] * e).sum((0, 1))
        dpre = ds[:, :, None] * v * (1 - e * e)
        dP += dpre
        dfh[t] = dpre.sum(1)
        dh = dh + dfh[t] @ Wa_h                                                                  # B4
    SN = S * N
    hp, hn = h_all[:S].reshape(SN, H), h_all[1:].reshape(SN, H)
    return dict(dP=dP, dM=dM, dv=dv, dwordtab=dwt, dWa_h=dfh.view(SN, H).t() @ hp, dW_ih=dgi.view(SN, -1).t() @ x_all.view(SN, X),
                db_ih=dgi.sum((0, 1)), dW_hh=dgh.view(SN, -1).t() @ hp, db_hh=dgh.sum((0, 1)),
                dW_out=dlogits.view(SN, V).t() @ hn, db_out=dlogits.sum((0, 1)))

