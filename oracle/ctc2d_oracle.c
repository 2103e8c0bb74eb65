/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported/linked by the product path
 * (megreader_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call into this file.
 *
 * CPU restatement of MegReader's 2D-CTC CUDA op.  The reference ships no CPU
 * implementation of this op (ops/ctc_2d/csrc/ctc2d.h:20,42 AT_ERROR on CPU) and
 * its CUDA sources do not build against torch 2.11 (THC/THC.h, AT_CHECK), so
 * this file follows the CUDA kernels statement by statement:
 *
 *   ctc2d_alpha   <- ops/ctc_2d/csrc/cuda/ctc2d_cuda_kernel.cu:54-211  (K1, log_alpha + nll)
 *                    allocation / initial values                :214-251 (at::zeros)
 *   ctc2d_beta    <- ops/ctc_2d/csrc/cuda/ctc2d_cuda_kernel.cu:254-368  (K2, log_beta)
 *   ctc2d_grad    <- ops/ctc_2d/csrc/cuda/ctc2d_cuda_kernel.cu:427-517  (K3, collect-all)
 *                    grad = full_like(log_probs, -inf)          :554
 *
 * Parity pin: the reference has no tests / golden vectors for this op
 * ("parity unpinned" by the reference itself).  It is pinned
 * instead against outputs of the reference's own pure-Python CTCLoss2D
 * (decoders/ctc_loss2d.py:86-154) generated in the build container by
 * oracle/make_golden.py and committed under tests/golden/.
 *
 * Layouts (all contiguous):
 *   log_probs [T,H,N,C]   targets [N,S] int64   input_lengths,target_lengths [N] int64
 *   log_alpha, log_beta [N,T,H,2S+1]   nll [N]   grad [T,H,N,C]
 *
 * Compiled twice: -DREAL=double -DSUF=f64 and -DREAL=float -DSUF=f32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#define SUF f64
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

static inline REAL r_exp(REAL x) { return sizeof(REAL) == 4 ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static inline REAL r_log(REAL x) { return sizeof(REAL) == 4 ? (REAL)logf((float)x) : (REAL)log((double)x); }

/* get_target_prime: ctc2d_cuda_kernel.cu:33-42 */
static inline int64_t target_prime(const int64_t *tg, int64_t S, int64_t b, int64_t idx, int64_t blank) {
    return (idx % 2 == 0) ? blank : tg[b * S + idx / 2];
}

/* safe_log_add: ctc2d_cuda_kernel.cu:44-51 */
static inline REAL safe_log_add(REAL a, REAL b) {
    REAL m = (a > b) ? a : b;
    if (m == -INFINITY) m = 0;
    return r_log(r_exp(a - m) + r_exp(b - m)) + m;
}

#define LP(t, h, b, c) lp[(((int64_t)(t) * H + (h)) * N + (b)) * C + (c)]
#define LA(b, t, h, s) la[(((int64_t)(b) * T + (t)) * H + (h)) * SS + (s)]
#define LB(b, t, h, s) lb[(((int64_t)(b) * T + (t)) * H + (h)) * SS + (s)]
#define GR(t, h, b, c) gr[(((int64_t)(t) * H + (h)) * N + (b)) * C + (c)]

/* K1.  la must hold N*T*H*(2S+1) values; it is zero-filled first like at::zeros (:230-233). */
void FN(ctc2d_alpha)(const REAL *lp, const int64_t *tg, const int64_t *in_len, const int64_t *tg_len,
                     int64_t T, int64_t H, int64_t N, int64_t C, int64_t S, int64_t blank,
                     REAL *la, REAL *nll) {
    const int64_t SS = 2 * S + 1;
    memset(la, 0, sizeof(REAL) * (size_t)(N * T * H * SS));
    for (int64_t b = 0; b < N; ++b) {
        const int64_t Tb = in_len[b], L = tg_len[b];
        /* t = 0  (:84-111) */
        for (int64_t s = 0; s < SS; ++s)
            for (int64_t h = 0; h < H; ++h) {
                REAL v;
                if (s == 0) v = LP(0, h, b, blank);
                else if (s == 1) v = (L > 0) ? LP(0, h, b, target_prime(tg, S, b, 1, blank)) : (REAL)-INFINITY;
                else v = (REAL)-INFINITY;
                LA(b, 0, h, s) = v;
            }
        /* recurrence (:125-184); every s reads row t-1 only, so a plain loop equals the barrier'd kernel */
        for (int64_t t = 1; t < T; ++t)
            for (int64_t s = 0; s < SS; ++s) {
                if (t < Tb && L > 0 && s < 2 * L + 1) {
                    const int64_t cur = target_prime(tg, S, b, s, blank);
                    const int have_three = (s > 1) && (target_prime(tg, S, b, s - 2, blank) != cur);
                    REAL la1 = LA(b, t - 1, 0, s);
                    for (int64_t h = 1; h < H; ++h) la1 = safe_log_add(la1, LA(b, t - 1, h, s));
                    REAL lamax = la1, la2, la3;
                    if (s > 0) {
                        la2 = LA(b, t - 1, 0, s - 1);
                        for (int64_t h = 1; h < H; ++h) la2 = safe_log_add(la2, LA(b, t - 1, h, s - 1));
                        if (la2 > lamax) lamax = la2;
                    } else la2 = (REAL)-INFINITY;
                    if (have_three) {
                        la3 = LA(b, t - 1, 0, s - 2);
                        for (int64_t h = 1; h < H; ++h) la3 = safe_log_add(la3, LA(b, t - 1, h, s - 2));
                        if (la3 > lamax) lamax = la3;
                    } else la3 = (REAL)-INFINITY;
                    if (lamax == -INFINITY) lamax = 0;
                    const REAL r = r_log(r_exp(la1 - lamax) + r_exp(la2 - lamax) + r_exp(la3 - lamax)) + lamax;
                    for (int64_t h = 0; h < H; ++h) LA(b, t, h, s) = r + LP(t, h, b, cur);
                } else {
                    for (int64_t h = 0; h < H; ++h) LA(b, t, h, s) = (REAL)-INFINITY;
                }
            }
        /* loss (:189-209).  L == 0 reads state -1 in the reference (undefined):
           here that term is taken as -inf. */
        {
            REAL l1 = LA(b, Tb - 1, 0, 2 * L);
            for (int64_t h = 1; h < H; ++h) l1 = safe_log_add(l1, LA(b, Tb - 1, h, 2 * L));
            REAL l2 = (REAL)-INFINITY;
            if (L > 0) {
                l2 = LA(b, Tb - 1, 0, 2 * L - 1);
                for (int64_t h = 1; h < H; ++h) l2 = safe_log_add(l2, LA(b, Tb - 1, h, 2 * L - 1));
            }
            REAL m = (l1 > l2) ? l1 : l2;
            if (m == -INFINITY) m = 0;
            nll[b] = -(r_log(r_exp(l1 - m) + r_exp(l2 - m)) + m);
        }
    }
}

/* K2.  lb zero-filled first like at::zeros (:536-539). */
void FN(ctc2d_beta)(const REAL *lp, const int64_t *tg, const int64_t *in_len, const int64_t *tg_len,
                    int64_t T, int64_t H, int64_t N, int64_t C, int64_t S, int64_t blank, REAL *lb) {
    const int64_t SS = 2 * S + 1;
    memset(lb, 0, sizeof(REAL) * (size_t)(N * T * H * SS));
    for (int64_t b = 0; b < N; ++b) {
        const int64_t Tb = in_len[b], L = tg_len[b];
        /* t = Tb-1 (:283-302) */
        for (int64_t s = 0; s < SS; ++s)
            for (int64_t h = 0; h < H; ++h) {
                REAL v;
                if (s == 2 * L) v = LP(Tb - 1, h, b, blank);
                else if (L > 0 && s == 2 * L - 1) v = LP(Tb - 1, h, b, target_prime(tg, S, b, s, blank));
                else v = (REAL)-INFINITY;
                LB(b, Tb - 1, h, s) = v;
            }
        for (int64_t t = T - 2; t >= 0; --t)
            for (int64_t s = 0; s < SS; ++s) {
                if (t < Tb - 1 && L > 0 && s < 2 * L + 1) {
                    const int64_t cur = target_prime(tg, S, b, s, blank);
                    const int have_three = (s < 2 * L - 1) && (target_prime(tg, S, b, s + 2, blank) != cur);
                    REAL lb1 = LB(b, t + 1, 0, s);
                    for (int64_t h = 1; h < H; ++h) lb1 = safe_log_add(lb1, LB(b, t + 1, h, s));
                    REAL lbmax = lb1, lb2, lb3;
                    if (s < 2 * L) {
                        lb2 = LB(b, t + 1, 0, s + 1);
                        for (int64_t h = 1; h < H; ++h) lb2 = safe_log_add(lb2, LB(b, t + 1, h, s + 1));
                        if (lb2 > lbmax) lbmax = lb2;
                    } else lb2 = (REAL)-INFINITY;
                    if (have_three) {
                        lb3 = LB(b, t + 1, 0, s + 2);
                        for (int64_t h = 1; h < H; ++h) lb3 = safe_log_add(lb3, LB(b, t + 1, h, s + 2));
                        if (lb3 > lbmax) lbmax = lb3;
                    } else lb3 = (REAL)-INFINITY;
                    if (lbmax == -INFINITY) lbmax = 0;
                    const REAL r = r_log(r_exp(lb1 - lbmax) + r_exp(lb2 - lbmax) + r_exp(lb3 - lbmax)) + lbmax;
                    for (int64_t h = 0; h < H; ++h) LB(b, t, h, s) = r + LP(t, h, b, cur);
                } else if (L == 0 || s > 2 * L + 1 || t >= Tb) {
                    for (int64_t h = 0; h < H; ++h) LB(b, t, h, s) = (REAL)-INFINITY;
                }
            }
    }
}

/* K3 (collect-all).  gr is filled with -inf first (:554). */
void FN(ctc2d_grad)(const REAL *grad_out, const REAL *lp, const int64_t *tg, const int64_t *in_len,
                    const int64_t *tg_len, const REAL *nll, const REAL *la, const REAL *lb,
                    int64_t T, int64_t H, int64_t N, int64_t C, int64_t S, int64_t blank, REAL *gr) {
    const int64_t SS = 2 * S + 1;
    for (int64_t i = 0; i < T * H * N * C; ++i) gr[i] = (REAL)-INFINITY;
    for (int64_t b = 0; b < N; ++b) {
        const int64_t Tb = in_len[b], L = tg_len[b];
        for (int64_t t = 0; t < T; ++t) {
            for (int64_t s = 0; s < SS; ++s) {
                if (L > 0 && s < 2 * L + 1) {
                    const int64_t cur = target_prime(tg, S, b, s, blank);
                    for (int64_t h = 0; h < H; ++h) {
                        const REAL lab = LA(b, t, h, s) + LB(b, t, h, s);
                        REAL *lcab = &GR(t, h, b, cur);
                        if (*lcab == -INFINITY) *lcab = lab;
                        else {
                            const REAL mx = (*lcab > lab) ? *lcab : lab;
                            *lcab = r_log(r_exp(*lcab - mx) + r_exp(lab - mx)) + mx;
                        }
                    }
                }
            }
            const REAL nl = nll[b], go = grad_out[b];
            for (int64_t c = 0; c < C; ++c)
                for (int64_t h = 0; h < H; ++h) {
                    REAL *res = &GR(t, h, b, c);
                    if (t < Tb) { /* zero_infinity hard-wired false (:526) */
                        const REAL l = LP(t, h, b, c);
                        if (*res == -INFINITY) *res = 0;
                        else *res = (r_exp(l) - r_exp(*res + nl - l)) * go;
                    } else *res = 0;
                }
        }
    }
}

/* forward + backward in one call, the unit bench.py's cpu_baseline times. */
void FN(ctc2d_fwd_bwd)(const REAL *grad_out, const REAL *lp, const int64_t *tg, const int64_t *in_len,
                       const int64_t *tg_len, int64_t T, int64_t H, int64_t N, int64_t C, int64_t S,
                       int64_t blank, REAL *nll, REAL *la, REAL *lb, REAL *gr) {
    FN(ctc2d_alpha)(lp, tg, in_len, tg_len, T, H, N, C, S, blank, la, nll);
    FN(ctc2d_beta)(lp, tg, in_len, tg_len, T, H, N, C, S, blank, lb);
    FN(ctc2d_grad)(grad_out, lp, tg, in_len, tg_len, nll, la, lb, T, H, N, C, S, blank, gr);
}
