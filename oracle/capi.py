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
