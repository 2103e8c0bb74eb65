"""Seeded synthetic inputs shared by the oracle tests, the GPU parity tests, smoke() and bench.py."""
import numpy as np


def ctc2d_case(seed, T, H, N, C, S, Lmax=None, peak=0.0, ragged_T=False, dtype=np.float32, blank=0):
    """SURVEY.md §8d synthetic 2D-CTC inputs: lp = log_softmax_H(mask)[...,None] + log_softmax_C(classify).
    Targets avoid the blank; feasibility (2L+1 <= T_b... loosely L <= T_b) is kept so nll is finite."""
    rng = np.random.RandomState(seed)
    Lmax = min(S, Lmax or S)
    il = np.full((N,), T, np.int64)
    if ragged_T:
        il = rng.randint(max(1, T // 2), T + 1, size=N).astype(np.int64)
    tl = np.empty((N,), np.int64)
    targets = np.zeros((N, S), np.int64)
    labels = [c for c in range(C) if c != blank]
    for b in range(N):
        lim = max(1, min(Lmax, int(il[b]) // 2))
        tl[b] = rng.randint(1, lim + 1)
        targets[b, :tl[b]] = rng.choice(labels, size=tl[b])
    m = rng.standard_normal((T, H, N)).astype(np.float64)
    c = rng.standard_normal((T, H, N, C)).astype(np.float64)
    if peak:
        for b in range(N):
            L = int(tl[b])
            for t in range(int(il[b])):
                k = min(L - 1, t * L // int(il[b]))
                c[t, :, b, targets[b, k]] += peak
    m = m - np.log(np.exp(m).sum(1, keepdims=True))
    c = c - np.log(np.exp(c).sum(3, keepdims=True))
    lp = (m[..., None] + c).astype(dtype)
    return np.ascontiguousarray(lp), targets, il, tl
