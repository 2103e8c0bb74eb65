"""ctypes loader for oracle/_build/liboracle.so (test infrastructure only).

Builds the library with `make -C oracle` on first use if it is missing (gcc is in the image).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in os.listdir(_HERE) if f.endswith(".c")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suf(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(dtype)


I64 = ctypes.c_int64


def ctc2d_forward(log_probs, targets, input_lengths, target_lengths, blank=0):
    """-> (nll[N], log_alpha[N,T,H,2S+1]); dtype follows log_probs (float32/float64)."""
    lp = np.ascontiguousarray(log_probs)
    T, H, N, C = lp.shape
    tg = np.ascontiguousarray(targets, dtype=np.int64)
    S = tg.shape[1]
    il = np.ascontiguousarray(input_lengths, dtype=np.int64)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int64)
    la = np.empty((N, T, H, 2 * S + 1), lp.dtype)
    nll = np.empty((N,), lp.dtype)
    fn = getattr(lib(), "ctc2d_alpha_" + _suf(lp.dtype))
    fn(ptr(lp), ptr(tg), ptr(il), ptr(tl), I64(T), I64(H), I64(N), I64(C), I64(S), I64(blank), ptr(la), ptr(nll))
    return nll, la


def ctc2d_beta(log_probs, targets, input_lengths, target_lengths, blank=0):
    lp = np.ascontiguousarray(log_probs)
    T, H, N, C = lp.shape
    tg = np.ascontiguousarray(targets, dtype=np.int64)
    S = tg.shape[1]
    il = np.ascontiguousarray(input_lengths, dtype=np.int64)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int64)
    lb = np.empty((N, T, H, 2 * S + 1), lp.dtype)
    fn = getattr(lib(), "ctc2d_beta_" + _suf(lp.dtype))
    fn(ptr(lp), ptr(tg), ptr(il), ptr(tl), I64(T), I64(H), I64(N), I64(C), I64(S), I64(blank), ptr(lb))
    return lb


def ctc2d_backward(grad_out, log_probs, targets, input_lengths, target_lengths, nll, log_alpha, blank=0):
    """-> grad[T,H,N,C] following K3 literally (computes log_beta internally with K2)."""
    lp = np.ascontiguousarray(log_probs)
    T, H, N, C = lp.shape
    tg = np.ascontiguousarray(targets, dtype=np.int64)
    S = tg.shape[1]
    il = np.ascontiguousarray(input_lengths, dtype=np.int64)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int64)
    go = np.ascontiguousarray(grad_out, dtype=lp.dtype)
    nl = np.ascontiguousarray(nll, dtype=lp.dtype)
    la = np.ascontiguousarray(log_alpha, dtype=lp.dtype)
    lb = ctc2d_beta(lp, tg, il, tl, blank)
    gr = np.empty_like(lp)
    fn = getattr(lib(), "ctc2d_grad_" + _suf(lp.dtype))
    fn(ptr(go), ptr(lp), ptr(tg), ptr(il), ptr(tl), ptr(nl), ptr(la), ptr(lb),
       I64(T), I64(H), I64(N), I64(C), I64(S), I64(blank), ptr(gr))
    return gr


def ctc2d_fwd_bwd(grad_out, log_probs, targets, input_lengths, target_lengths, blank=0):
    """One call: (nll, log_alpha, grad).  The unit bench.py's cpu_baseline times."""
    lp = np.ascontiguousarray(log_probs)
    T, H, N, C = lp.shape
    tg = np.ascontiguousarray(targets, dtype=np.int64)
    S = tg.shape[1]
    il = np.ascontiguousarray(input_lengths, dtype=np.int64)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int64)
    go = np.ascontiguousarray(grad_out, dtype=lp.dtype)
    la = np.empty((N, T, H, 2 * S + 1), lp.dtype)
    lb = np.empty_like(la)
    nll = np.empty((N,), lp.dtype)
    gr = np.empty_like(lp)
    fn = getattr(lib(), "ctc2d_fwd_bwd_" + _suf(lp.dtype))
    fn(ptr(go), ptr(lp), ptr(tg), ptr(il), ptr(tl), I64(T), I64(H), I64(N), I64(C), I64(S), I64(blank),
       ptr(nll), ptr(la), ptr(lb), ptr(gr))
    return nll, la, gr


# ---------------------------------------------------------------- DCN (oracle/dcn_oracle.c)
I32 = ctypes.c_int


def _dcn_out(H, W, kh, kw, sh, sw, ph, pw, dh, dw):
    return ((H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1, (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1)


def dcn_forward(inp, weight, bias, offset, mask, stride=1, padding=1, dilation=1, group=1, dg=1):
    """Modulated (mask given) or v1 (mask None) deformable conv forward.  offset/mask may have a larger spatial
    size than the output: they are indexed flat per sample with (Ho, Wo) strides like the reference kernels."""
    x = np.ascontiguousarray(inp)
    dt = x.dtype
    B, C, H, W = x.shape
    w = np.ascontiguousarray(weight, dtype=dt)
    Cout, _, kh, kw = w.shape
    off = np.ascontiguousarray(offset, dtype=dt)
    msk = None if mask is None else np.ascontiguousarray(mask, dtype=dt)
    Ho, Wo = _dcn_out(H, W, kh, kw, stride, stride, padding, padding, dilation, dilation)
    out = np.empty((B, Cout, Ho, Wo), dt)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=dt)
    fn = getattr(lib(), "dcn_forward_" + _suf(dt))
    fn(ptr(x), ptr(w), ptr(b) if b is not None else None, ptr(off), I64(off[0].size),
       ptr(msk) if msk is not None else None, I64(msk[0].size if msk is not None else 0),
       I32(B), I32(C), I32(H), I32(W), I32(Cout), I32(kh), I32(kw), I32(stride), I32(stride), I32(padding),
       I32(padding), I32(dilation), I32(dilation), I32(group), I32(dg), I32(1 if b is not None else 0), ptr(out))
    return out


def dcn_backward(inp, weight, bias, offset, mask, grad_output, stride=1, padding=1, dilation=1, group=1, dg=1):
    """-> (grad_input, grad_weight, grad_bias|None, grad_offset, grad_mask|None); grad_offset/grad_mask have the
    shape of offset/mask with the reference's flat (Ho,Wo) layout inside each sample slab (tail left zero)."""
    x = np.ascontiguousarray(inp)
    dt = x.dtype
    B, C, H, W = x.shape
    w = np.ascontiguousarray(weight, dtype=dt)
    Cout, _, kh, kw = w.shape
    off = np.ascontiguousarray(offset, dtype=dt)
    msk = None if mask is None else np.ascontiguousarray(mask, dtype=dt)
    go = np.ascontiguousarray(grad_output, dtype=dt)
    gi, gw = np.zeros_like(x), np.zeros_like(w)
    gb = None if bias is None else np.zeros((Cout,), dt)
    goff = np.zeros_like(off)
    gmsk = None if msk is None else np.zeros_like(msk)
    fn = getattr(lib(), "dcn_backward_" + _suf(dt))
    fn(ptr(x), ptr(w), ptr(off), I64(off[0].size), ptr(msk) if msk is not None else None,
       I64(msk[0].size if msk is not None else 0), ptr(go),
       I32(B), I32(C), I32(H), I32(W), I32(Cout), I32(kh), I32(kw), I32(stride), I32(stride), I32(padding),
       I32(padding), I32(dilation), I32(dilation), I32(group), I32(dg), I32(1 if gb is not None else 0),
       ptr(gi), ptr(gw), ptr(gb) if gb is not None else None, ptr(goff), I64(goff[0].size),
       ptr(gmsk) if gmsk is not None else None, I64(gmsk[0].size if gmsk is not None else 0))
    return gi, gw, gb, goff, gmsk


# ------------------------------------------------------------------ deformable PS-RoI pooling (oracle/deform_pool_oracle.c)
def deform_psroi_forward(data, rois, trans, no_trans, spatial_scale, output_dim, group_size, pooled, part_size,
                         sample_per_part, trans_std):
    x = np.ascontiguousarray(data)
    dt = x.dtype
    B, C, H, W = x.shape
    r = np.ascontiguousarray(rois, dtype=dt)
    n = r.shape[0]
    t = None if no_trans else np.ascontiguousarray(trans, dtype=dt)
    ct = 2 if no_trans else t.shape[1]
    out = np.zeros((n, output_dim, pooled, pooled), dt)
    cnt = np.zeros_like(out)
    real = ctypes.c_double if dt == np.float64 else ctypes.c_float
    fn = getattr(lib(), "deform_psroi_forward_" + _suf(dt))
    fn(ptr(x), ptr(r), ptr(t) if t is not None else None, I32(C), I32(H), I32(W), I32(n), I32(ct), I32(int(no_trans)),
       real(spatial_scale), I32(output_dim), I32(group_size), I32(pooled), I32(part_size), I32(sample_per_part),
       real(trans_std), ptr(out), ptr(cnt))
    return out, cnt


def deform_psroi_backward(out_grad, data, rois, trans, top_count, no_trans, spatial_scale, output_dim, group_size, pooled,
                          part_size, sample_per_part, trans_std):
    x = np.ascontiguousarray(data)
    dt = x.dtype
    B, C, H, W = x.shape
    r = np.ascontiguousarray(rois, dtype=dt)
    n = r.shape[0]
    t = None if no_trans else np.ascontiguousarray(trans, dtype=dt)
    ct = 2 if no_trans else t.shape[1]
    go = np.ascontiguousarray(out_grad, dtype=dt)
    tc = np.ascontiguousarray(top_count, dtype=dt)
    gin = np.zeros_like(x)
    gtr = None if no_trans else np.zeros_like(t)
    real = ctypes.c_double if dt == np.float64 else ctypes.c_float
    fn = getattr(lib(), "deform_psroi_backward_" + _suf(dt))
    fn(ptr(go), ptr(x), ptr(r), ptr(t) if t is not None else None, ptr(tc), I32(C), I32(H), I32(W), I32(n), I32(ct),
       I32(int(no_trans)), real(spatial_scale), I32(output_dim), I32(group_size), I32(pooled), I32(part_size),
       I32(sample_per_part), real(trans_std), ptr(gin), ptr(gtr) if gtr is not None else None)
    return gin, gtr
