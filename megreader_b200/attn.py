"""Host side of the attention recogniser head's recurrent loop (csrc/attn_decode.cu).

    attention_loop_loss(projected, memory_bt, cell, targets, lengths, feedback, blank)  ->  loss (N,), attention (N, S, L)

is the training branch of decoders/attention_decoder.py:96-117 for everything after the encoder: the max_size steps of
AttentionRNNCell.forward (:187-231) with the reference's teacher forcing / step dropout, the masked NLL summed over the steps and the
attention maps -- ONE persistent cooperative kernel forward, one backward (through time), plus four dense weight-gradient products
over the S*N saved rows (plain library GEMMs).  fp32 like the parameters.  CUDA only: there is no CPU fallback.
"""
import ctypes

import torch

from . import _lib


def _f32(t):
    return t.detach().float().contiguous()


def _i32(t, dev):
    return t.detach().to(device=dev, dtype=torch.int32).contiguous()


def _check_sync(sync, dev, what):
    """read the kernels' error word unless a CUDA graph is being captured (then the caller checks `sync` after the replay)"""
    if torch.cuda.is_current_stream_capturing():
        return
    status = ctypes.c_int(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mr_attn_sync_status(sync.data_ptr(), torch.cuda.current_stream().cuda_stream, ctypes.byref(status)),
                   "attn_sync_status")
    if status.value:
        raise RuntimeError("megreader_b200 %s: grid barrier timed out (error word %d)" % (what, status.value))


class _AttnLoopFn(torch.autograd.Function):
    """inputs: projected (N,L,H), memory (N,L,H+E), wa_h (H,H) [= attn.attn.weight[:, :H], any row stride], v (H), wordtab (V,H),
    w_ih (3H,2H+E), b_ih, w_hh (3H,H), b_hh, w_out (V,H), b_out, then int32 targets (N,S), lengths (N), coin (S), swap / noise (S,N)."""

    @staticmethod
    def forward(ctx, projected, memory, wa_h, v, wordtab, w_ih, b_ih, w_hh, b_hh, w_out, b_out, targets, lengths, coin, swap, noise,
                blank, check):
        if not projected.is_cuda:
            raise NotImplementedError("megreader_b200.attn: CUDA tensors only (no CPU fallback)")
        dev = projected.device
        N, L, H = projected.shape
        D = memory.shape[2]
        E = D - H
        V = w_out.shape[0]
        S = targets.shape[1]
        X = H + D
        assert memory.shape == (N, L, D) and wa_h.shape == (H, H) and wordtab.shape == (V, H) and w_ih.shape == (3 * H, X)
        assert w_hh.shape == (3 * H, H) and targets.shape == (N, S) and coin.shape == (S,) and swap.shape == (S, N) == noise.shape
        projected, memory, v, wordtab = _f32(projected), _f32(memory), _f32(v), _f32(wordtab)
        w_ih, b_ih, w_hh, b_hh, w_out, b_out = (_f32(t) for t in (w_ih, b_ih, w_hh, b_hh, w_out, b_out))
        wa = wa_h.detach().float()
        if wa.stride(1) != 1:
            wa = wa.contiguous()
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)  # noqa: E731
        Xp = -(-X // 4) * 4                                   # GRU-input rows padded to 16 bytes (cp.async staging in the kernel)
        h_all, fh_all, x_all, gates = f(S + 1, N, H), f(S, N, H), f(S, N, Xp), f(S, N, 4, H)
        logp, attn, loss = f(S, N, V), f(N, S, L), f(N)
        word = torch.empty((S, N), dtype=torch.int32, device=dev)
        sync = torch.empty(2, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mr_attn_train_fwd_f32(
                projected.data_ptr(), memory.data_ptr(), wa.data_ptr(), wa.stride(0), v.data_ptr(), wordtab.data_ptr(),
                w_ih.data_ptr(), b_ih.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(), w_out.data_ptr(), b_out.data_ptr(),
                targets.data_ptr(), lengths.data_ptr(), coin.data_ptr(), swap.data_ptr(), noise.data_ptr(),
                h_all.data_ptr(), fh_all.data_ptr(), x_all.data_ptr(), gates.data_ptr(), logp.data_ptr(), attn.data_ptr(),
                word.data_ptr(), loss.data_ptr(), sync.data_ptr(), N, L, H, E, V, S, int(blank),
                torch.cuda.current_stream().cuda_stream), "attn_train_fwd")
        if check:
            _check_sync(sync, dev, "attn_train_fwd")
        ctx.save_for_backward(projected, memory, wa, v, w_ih, w_hh, w_out, h_all, fh_all, x_all, gates, logp, attn, word, targets, lengths)
        ctx.dims = (N, L, H, E, V, S)
        ctx.check = check
        ctx.mark_non_differentiable(attn, word)
        return loss, attn, word

    @staticmethod
    def backward(ctx, grad_loss, _grad_attn, _grad_word):
        projected, memory, wa, v, w_ih, w_hh, w_out, h_all, fh_all, x_all, gates, logp, attn, word, targets, lengths = ctx.saved_tensors
        N, L, H, E, V, S = ctx.dims
        D, X = H + E, 2 * H + E
        dev = projected.device
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)  # noqa: E731
        dlogits, dgi, dgh, dfh = f(S, N, V), f(S, N, 3 * H), f(S, N, 3 * H), f(S, N, H)
        dx, dh, dP, dM, dv, dwt = f(N, X), f(N, H), f(N, L, H), f(N, L, D), f(H), f(V, H)
        sync = torch.empty(2, dtype=torch.int32, device=dev)
        g = grad_loss.detach().float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mr_attn_train_bwd_f32(
                projected.data_ptr(), memory.data_ptr(), wa.data_ptr(), wa.stride(0), v.data_ptr(), w_ih.data_ptr(), w_hh.data_ptr(),
                w_out.data_ptr(), h_all.data_ptr(), fh_all.data_ptr(), gates.data_ptr(), logp.data_ptr(), attn.data_ptr(),
                word.data_ptr(), targets.data_ptr(), lengths.data_ptr(), g.data_ptr(), dlogits.data_ptr(), dgi.data_ptr(),
                dgh.data_ptr(), dfh.data_ptr(), dx.data_ptr(), dh.data_ptr(), dP.data_ptr(), dM.data_ptr(), dv.data_ptr(),
                dwt.data_ptr(), sync.data_ptr(), N, L, H, E, V, S, torch.cuda.current_stream().cuda_stream), "attn_train_bwd")
        if ctx.check:
            _check_sync(sync, dev, "attn_train_bwd")
        # weight gradients: dense products over the S*N saved rows (plain library GEMMs, fp32)
        h_prev, h_next = h_all[:S].reshape(S * N, H), h_all[1:].reshape(S * N, H)
        dl2, gi2, gh2, fh2 = dlogits.view(S * N, V), dgi.view(S * N, 3 * H), dgh.view(S * N, 3 * H), dfh.view(S * N, H)
        need = ctx.needs_input_grad
        d_wa = fh2.t().mm(h_prev) if need[2] else None
        d_wih = gi2.t().mm(x_all.view(S * N, -1)[:, :X]) if need[5] else None
        d_bih = gi2.sum(0) if need[6] else None
        d_whh = gh2.t().mm(h_prev) if need[7] else None
        d_bhh = gh2.sum(0) if need[8] else None
        d_wout = dl2.t().mm(h_next) if need[9] else None
        d_bout = dl2.sum(0) if need[10] else None
        return (dP if need[0] else None, dM if need[1] else None, d_wa, dv if need[3] else None, dwt if need[4] else None,
                d_wih, d_bih, d_whh, d_bhh, d_wout, d_bout, None, None, None, None, None, None, None)


def attention_loop_loss(projected, memory_bt, cell, targets, lengths, feedback, blank, check=True, return_words=False):
    """projected (N,L,H) = cell.attn.project_encoder(memory), memory_bt (N,L,H+E), cell = AttentionRNNCell, targets (N,>=S) integer,
    lengths (N,), feedback = (coin (S,) bool, swap (S,N), noise (S,N)) from AttentionDecoder.draw_feedback
    -> loss (N,) fp32 = sum_t NLL_t * (t <= lengths), attention (N,S,L) [, words (S,N) int32: the symbol fed into every step]."""
    dev = projected.device
    coin, swap, noise = feedback
    S = coin.shape[0]
    H = cell.hidden_dims
    wordtab = cell.word_linear(cell.embedding.weight)                        # (V,H): row w = word_linear(embedding(w)); autograd-visible
    loss, maps, words = _AttnLoopFn.apply(
        projected, memory_bt, cell.attn.attn.weight[:, :H], cell.attn.v, wordtab, cell.rnn.weight_ih, cell.rnn.bias_ih,
        cell.rnn.weight_hh, cell.rnn.bias_hh, cell.out.weight, cell.out.bias, _i32(targets[:, :S], dev), _i32(lengths, dev),
        _i32(coin, dev), _i32(swap, dev), _i32(noise, dev), int(blank), bool(check))
    return (loss, maps, words) if return_words else (loss, maps)
