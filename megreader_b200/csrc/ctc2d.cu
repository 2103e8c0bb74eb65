// 2D-CTC loss for sm_100a — forward (log_alpha + nll), backward (gradient), and the training pair
// (forward_train + backward_apply).  Replaces the reference's four SIMT kernels
// (ops/ctc_2d/csrc/cuda/ctc2d_cuda_kernel.cu:54-211 K1, :254-368 K2, :427-517 K3) — not a port:
//
// Observation that shapes the design.  In the reference recurrence every height shares one
// transition term:  alpha[t,h,s] = lp[t,h,l'_s] + R[t,s]  with
//     R[t,s]   = LSE( A[t-1,s], A[t-1,s-1], [l'_s != l'_{s-2}] A[t-1,s-2] )
//     A[t,s]   = LSE_h alpha[t,h,s] = R[t,s] + Q[t,l'_s],      Q[t,c] = LSE_h lp[t,h,c]
// so the dynamic programme is a 1D CTC over the height-marginal Q (tiny, on chip), and log_alpha is a
// pure streaming expansion of lp.  Likewise beta[t,h,s] = lp[t,h,l'_s] + Rb[t,s], and K3's gradient
//     grad[t,h,c] = (exp(lp) - exp(LSE_{s:l'_s=c}(alpha+beta) + nll - lp)) * go
//                 = exp(lp[t,h,c]) * (1 - sum_{s:l'_s=c} exp(R[t,s] + Rb[t,s] + nll)) * go      (class present)
// needs only a per-(t,c) factor.  Hence: log_probs is read ONCE per kernel, log_beta never exists in
// HBM, and the backward never reads log_alpha.
//
// Kernels (all HBM-bound; algorithmic bytes per sample, fp32, cfg-3 shape T32 H8 C38 S32):
//   ctc2d_alpha_kernel   : read |lp| 38,912 + write |alpha| 66,560            (contract forward)
//   ctc2d_dp_kernel<GRAD>: read |lp| (+ L2 re-read) + write |grad| 38,912      (contract backward)
//   ctc2d_dp_kernel<FAC> : read |lp| + write gfac T*C*4 = 4,864                (training forward)
//   ctc2d_apply_kernel   : read |lp| + gfac, write |grad|                      (training backward)
#include "common.cuh"
#include <math.h>
#include <string.h>

namespace {

using namespace mr;

constexpr int kStages = 4;  // cp.async ring depth of the alpha kernel

template <typename real> struct Lim;
template <> struct Lim<float> { static __device__ __forceinline__ float ninf() { return -INFINITY; } };
template <> struct Lim<double> { static __device__ __forceinline__ double ninf() { return -(double)INFINITY; } };

template <bool FAST> __device__ __forceinline__ float ex(float x) { return FAST ? __expf(x) : expf(x); }
template <bool FAST> __device__ __forceinline__ float lg(float x) { return FAST ? __logf(x) : logf(x); }
template <bool FAST> __device__ __forceinline__ double ex(double x) { return exp(x); }
template <bool FAST> __device__ __forceinline__ double lg(double x) { return log(x); }

// LSE with the all -inf case returning -inf (safe_log_add, ctc2d_cuda_kernel.cu:44-51 / :166-167)
template <bool FAST, typename real>
__device__ __forceinline__ real lse2(real a, real b) {
    real m = fmax(a, b);
    if (m == Lim<real>::ninf()) return m;
    return m + lg<FAST>(ex<FAST>(a - m) + ex<FAST>(b - m));
}
template <bool FAST, typename real>
__device__ __forceinline__ real lse3(real a, real b, real c) {
    real m = fmax(a, fmax(b, c));
    if (m == Lim<real>::ninf()) return m;
    return m + lg<FAST>(ex<FAST>(a - m) + ex<FAST>(b - m) + ex<FAST>(c - m));
}

struct Geo {
    int T, H, N, C, S, SS;  // SS = 2S+1
    int G;                  // samples per CTA
    int vec;                // elements per 16-byte vector usable on log_probs rows (1 = scalar path)
    int blank;
    int64_t tg_sn, tg_ss;
};

__device__ __forceinline__ int clampi(int64_t v, int hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : (int)v); }

// Per-thread view of one (sample, state): lengths, l'_s and the skip flags (K1 :113-123, K2 :304-314).
struct StateCtx {
    bool active;
    int g, s, b, cur;
    int64_t Tb, L;
    bool skip_fwd, skip_bwd, in_range;  // in_range: L > 0 && s <= 2L
};
__device__ __forceinline__ StateCtx make_ctx(const Geo &q, int b0, int Gv, const int64_t *tg, const int64_t *il,
                                             const int64_t *tl) {
    StateCtx c;
    const int tid = threadIdx.x;
    c.g = tid / q.SS;
    c.s = tid - c.g * q.SS;
    c.active = (tid < q.G * q.SS) && (c.g < Gv);
    c.b = b0 + c.g;
    c.cur = q.blank;
    c.Tb = 0; c.L = 0;
    c.skip_fwd = c.skip_bwd = c.in_range = false;
    if (c.active) {
        c.Tb = il[c.b];
        c.L = tl[c.b];
        if (c.s < 2 * c.L + 1) {
            c.in_range = c.L > 0;
            const int64_t *row = tg + (int64_t)c.b * q.tg_sn;
            if (c.s & 1) {
                const int64_t me = row[(int64_t)(c.s >> 1) * q.tg_ss];
                c.cur = clampi(me, q.C);
                if (c.s > 1) c.skip_fwd = row[(int64_t)((c.s - 2) >> 1) * q.tg_ss] != me;
                if (c.s < 2 * c.L - 1) c.skip_bwd = row[(int64_t)((c.s + 2) >> 1) * q.tg_ss] != me;
            }
        }
    }
    return c;
}

// ------------------------------------------------------------------------------------------------
// Contract forward: log_alpha + nll.  One CTA = G consecutive samples, thread = (sample, state).
// The [H, G*C] slab of log_probs for column t is prefetched kStages-1 columns ahead with cp.async;
// it feeds both Q[t] and the expansion, so HBM sees every byte of log_probs exactly once.
// ------------------------------------------------------------------------------------------------
template <typename real, bool STAGED>
__device__ __forceinline__ void load_slab(const Geo &q, const real *__restrict__ lp, real *dst, int t, int b0, int Gv) {
    if (!STAGED) return;
    const int rowElems = q.G * q.C;
    const int n_el = Gv * q.C;
    const int nvec = n_el / q.vec;
    const int rem = n_el - nvec * q.vec;
    const int per_row = nvec + rem;
    for (int i = threadIdx.x; i < q.H * per_row; i += blockDim.x) {
        const int h = i / per_row, j = i - h * per_row;
        const real *src = lp + ((int64_t)(t * q.H + h) * q.N + b0) * q.C;
        real *d = dst + h * rowElems;
        if (j < nvec) {
            if (q.vec > 1) cp_async16(d + j * q.vec, src + j * q.vec);
            else if (sizeof(real) == 8) cp_async8(d + j, src + j);
            else cp_async4(d + j, src + j);
        } else {
            const int e = nvec * q.vec + (j - nvec);
            if (sizeof(real) == 8) cp_async8(d + e, src + e);
            else cp_async4(d + e, src + e);
        }
    }
}

template <typename real, bool FAST, bool STAGED>
__global__ void ctc2d_alpha_kernel(Geo q, const real *__restrict__ lp, const int64_t *__restrict__ tg,
                                   const int64_t *__restrict__ il, const int64_t *__restrict__ tl,
                                   real *__restrict__ nll, real *__restrict__ la) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const real NINF = Lim<real>::ninf();
    const int tid = threadIdx.x, nth = blockDim.x;
    const int b0 = blockIdx.x * q.G;
    const int Gv = min(q.G, q.N - b0);
    const int rowElems = q.G * q.C;
    const int slabElems = q.H * rowElems;
    real *slab = reinterpret_cast<real *>(smem_raw);
    real *Qs = slab + (STAGED ? kStages * slabElems : 0);
    real *As = Qs + rowElems;
    real *fin = As + q.G * q.SS;

    const StateCtx c = make_ctx(q, b0, Gv, tg, il, tl);
    if (tid < 2 * q.G) fin[tid] = NINF;
    real R = NINF;

    if (STAGED) {
        for (int st = 0; st < kStages - 1; ++st) {
            if (st < q.T) load_slab<real, STAGED>(q, lp, slab + st * slabElems, st, b0, Gv);
            cp_async_commit();
        }
    }
    for (int t = 0; t < q.T; ++t) {
        if (STAGED) cp_async_wait<kStages - 2>();
        __syncthreads();  // slab[t] landed; As/fin of step t-1 visible; slab[t-1] free for reuse
        if (STAGED) {
            const int tn = t + kStages - 1;
            if (tn < q.T) load_slab<real, STAGED>(q, lp, slab + (tn % kStages) * slabElems, tn, b0, Gv);
            cp_async_commit();
        }
        const real *sl = STAGED ? slab + (t % kStages) * slabElems : lp + ((int64_t)t * q.H * q.N + b0) * q.C;
        const int64_t hstride = STAGED ? rowElems : (int64_t)q.N * q.C;

        if (c.active) {
            if (t == 0) {
                R = (c.s == 0 || (c.s == 1 && c.L > 0)) ? (real)0 : NINF;  // K1 :84-111
            } else if (t < c.Tb && c.in_range) {                           // K1 :128-173
                const real *a = As + c.g * q.SS + c.s;
                const real a1 = a[0];
                const real a2 = c.s > 0 ? a[-1] : NINF;
                const real a3 = c.skip_fwd ? a[-2] : NINF;
                R = lse3<FAST>(a1, a2, a3);
            } else {
                R = NINF;                                                  // K1 :174-182
            }
            real *out = la + (((int64_t)c.b * q.T + t) * q.H) * q.SS + c.s;
            const real *src = sl + c.g * q.C + c.cur;
#pragma unroll 4
            for (int h = 0; h < q.H; ++h) __stcs(out + (int64_t)h * q.SS, src[h * hstride] + R);
        }
        // Q[t][g][c] = LSE_h lp[t,h,b,c]
        for (int idx = tid; idx < Gv * q.C; idx += nth) {
            const real *p = sl + idx;
            real m = NINF;
            for (int h = 0; h < q.H; ++h) m = fmax(m, p[h * hstride]);
            real v = m;
            if (m != NINF) {
                real sum = 0;
                for (int h = 0; h < q.H; ++h) sum += ex<FAST>(p[h * hstride] - m);
                v = m + lg<FAST>(sum);
            }
            Qs[idx] = v;
        }
        __syncthreads();
        if (c.active) {
            const real a = R + Qs[c.g * q.C + c.cur];
            As[c.g * q.SS + c.s] = a;
            if (t == c.Tb - 1) {  // K1 :189-209 reads LSE_h alpha[Tb-1, h, 2L] and [.., 2L-1]
                if (c.s == 2 * c.L) fin[2 * c.g] = a;
                else if (c.s == 2 * c.L - 1) fin[2 * c.g + 1] = a;
            }
        }
    }
    __syncthreads();
    if (c.active && c.s == 0) nll[c.b] = -lse2<FAST>(fin[2 * c.g], fin[2 * c.g + 1]);
}

// ------------------------------------------------------------------------------------------------
// DP kernel (contract backward / training forward).
//   P1  stream log_probs, Q[t][g][c] for all t into shared memory
//   P2  forward sweep R[t] (kept in smem), nll; backward sweep Rb[t] fused with the per-class
//       accumulation  acc[t][g][c] += exp(R + Rb + nll)   and the "class present" flag
//   P3  MODE_GRAD: re-stream log_probs (L2) and write grad = exp(lp) * fac * go
//       MODE_FAC : write fac [N,T,C] (+ nll)
// ------------------------------------------------------------------------------------------------
enum { MODE_GRAD = 0, MODE_FAC = 1 };

template <typename real, int VE> struct alignas(sizeof(real) * VE) VecT { real v[VE]; };

// 4/8/16-byte vector loads/stores through the builtin types (so the cache-hint intrinsics apply)
template <int BYTES> struct Raw;
template <> struct Raw<16> { using type = float4; };
template <> struct Raw<8> { using type = float2; };
template <> struct Raw<4> { using type = float; };
template <typename real, int VE>
__device__ __forceinline__ VecT<real, VE> ld_nc(const real *p) {
    using R = typename Raw<sizeof(real) * VE>::type;
    const R t = __ldg(reinterpret_cast<const R *>(p));
    VecT<real, VE> r;
    memcpy(&r, &t, sizeof(R));
    return r;
}
template <typename real, int VE>
__device__ __forceinline__ VecT<real, VE> ld_cs(const real *p) {
    using R = typename Raw<sizeof(real) * VE>::type;
    const R t = __ldcs(reinterpret_cast<const R *>(p));
    VecT<real, VE> r;
    memcpy(&r, &t, sizeof(R));
    return r;
}
template <typename real, int VE>
__device__ __forceinline__ void st_cs(real *p, const VecT<real, VE> &v) {
    using R = typename Raw<sizeof(real) * VE>::type;
    R t;
    memcpy(&t, &v, sizeof(R));
    __stcs(reinterpret_cast<R *>(p), t);
}

template <typename real, bool FAST, int VE>
__device__ __forceinline__ void phase_q(const Geo &q, const real *__restrict__ lp, real *Qall, int b0, int Gv) {
    const int rowElems = q.G * q.C;
    const int n_el = Gv * q.C;
    const int nvec = (n_el + VE - 1) / VE;  // VE > 1 only when n_el % VE == 0 for every CTA
    const real NINF = Lim<real>::ninf();
    for (int i = threadIdx.x; i < q.T * nvec; i += blockDim.x) {
        const int t = i / nvec, j = i - t * nvec;
        const real *base = lp + ((int64_t)t * q.H * q.N + b0) * q.C + j * VE;
        const int64_t hs = (int64_t)q.N * q.C;
        real m[VE], sum[VE];
#pragma unroll
        for (int k = 0; k < VE; ++k) { m[k] = NINF; sum[k] = 0; }
        for (int h0 = 0; h0 < q.H; h0 += 8) {
            VecT<real, VE> x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (h0 + u < q.H) x[u] = ld_nc<real, VE>(base + (h0 + u) * hs);
                else {
#pragma unroll
                    for (int k = 0; k < VE; ++k) x[u].v[k] = NINF;
                }
            }
#pragma unroll
            for (int k = 0; k < VE; ++k) {
                real cm = x[0].v[k];
#pragma unroll
                for (int u = 1; u < 8; ++u) cm = fmax(cm, x[u].v[k]);
                const real nm = fmax(m[k], cm);
                if (nm != NINF) {
                    real sacc = (m[k] == NINF) ? (real)0 : sum[k] * ex<FAST>(m[k] - nm);
#pragma unroll
                    for (int u = 0; u < 8; ++u) sacc += ex<FAST>(x[u].v[k] - nm);
                    sum[k] = sacc;
                    m[k] = nm;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < VE; ++k)
            Qall[t * rowElems + j * VE + k] = (m[k] == NINF) ? NINF : m[k] + lg<FAST>(sum[k]);
    }
}

template <typename real, bool FAST, int VE>
__device__ __forceinline__ void phase_grad(const Geo &q, const real *__restrict__ lp, const real *Fs,
                                           real *__restrict__ grad, int b0, int Gv) {
    const int rowElems = q.G * q.C;
    const int n_el = Gv * q.C;
    const int nvec = (n_el + VE - 1) / VE;
    const int64_t rows = (int64_t)q.T * q.H;
    for (int64_t i = threadIdx.x; i < rows * nvec; i += blockDim.x) {
        const int64_t r = i / nvec;
        const int j = (int)(i - r * nvec);
        const int t = (int)(r / q.H);
        const int64_t off = (r * q.N + b0) * q.C + j * VE;
        VecT<real, VE> x = ld_cs<real, VE>(lp + off);
        const real *f = Fs + t * rowElems + j * VE;
        VecT<real, VE> o;
#pragma unroll
        for (int k = 0; k < VE; ++k) {
            const real fk = f[k];
            o.v[k] = (fk == (real)0) ? (real)0 : ex<FAST>(x.v[k]) * fk;
        }
        st_cs<real, VE>(grad + off, o);
    }
}

template <typename real, bool FAST, int MODE>
__global__ void ctc2d_dp_kernel(Geo q, const real *__restrict__ lp, const int64_t *__restrict__ tg,
                                const int64_t *__restrict__ il, const int64_t *__restrict__ tl,
                                const real *__restrict__ grad_out, int64_t go_stride,
                                real *__restrict__ nll_out, real *__restrict__ fac_out, real *__restrict__ grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const real NINF = Lim<real>::ninf();
    const int tid = threadIdx.x, nth = blockDim.x;
    const int b0 = blockIdx.x * q.G;
    const int Gv = min(q.G, q.N - b0);
    const int rowElems = q.G * q.C;
    const int rowStates = q.G * q.SS;
    real *Qall = reinterpret_cast<real *>(smem_raw);            // [T][G*C]
    real *acc = Qall + q.T * rowElems;                           // [T][G*C]  sum of exp(R+Rb+nll), then factor
    real *Ra = acc + q.T * rowElems;                             // [T][G*SS]
    real *As = Ra + q.T * rowStates;                             // [2][G*SS]
    real *fin = As + 2 * rowStates;                              // [2G] final states, then [G] nll
    unsigned char *pres = reinterpret_cast<unsigned char *>(fin + 3 * q.G);  // [T][G*C]

    const StateCtx c = make_ctx(q, b0, Gv, tg, il, tl);
    for (int i = tid; i < q.T * rowElems; i += nth) { acc[i] = 0; pres[i] = 0; }
    if (tid < 2 * q.G) fin[tid] = NINF;

    if (q.vec == 4) phase_q<real, FAST, (sizeof(real) == 4 ? 4 : 1)>(q, lp, Qall, b0, Gv);
    else if (q.vec == 2) phase_q<real, FAST, (sizeof(real) == 8 ? 2 : 1)>(q, lp, Qall, b0, Gv);
    else phase_q<real, FAST, 1>(q, lp, Qall, b0, Gv);
    __syncthreads();

    // ---- forward sweep (same recurrence as ctc2d_alpha_kernel), one barrier per column
    real R = NINF;
    for (int t = 0; t < q.T; ++t) {
        if (c.active) {
            if (t == 0) R = (c.s == 0 || (c.s == 1 && c.L > 0)) ? (real)0 : NINF;
            else if (t < c.Tb && c.in_range) {
                const real *a = As + ((t - 1) & 1) * rowStates + c.g * q.SS + c.s;
                R = lse3<FAST>(a[0], c.s > 0 ? a[-1] : NINF, c.skip_fwd ? a[-2] : NINF);
            } else R = NINF;
            Ra[t * rowStates + c.g * q.SS + c.s] = R;
            const real a = R + Qall[t * rowElems + c.g * q.C + c.cur];
            As[(t & 1) * rowStates + c.g * q.SS + c.s] = a;
            if (t == c.Tb - 1) {
                if (c.s == 2 * c.L) fin[2 * c.g] = a;
                else if (c.s == 2 * c.L - 1) fin[2 * c.g + 1] = a;
            }
        }
        __syncthreads();
    }
    real *nlls = fin + 2 * q.G;
    if (c.active && c.s == 0) {
        const real v = -lse2<FAST>(fin[2 * c.g], fin[2 * c.g + 1]);
        nlls[c.g] = v;
        if (MODE == MODE_FAC) nll_out[c.b] = v;
    }
    __syncthreads();
    const real my_nll = c.active ? nlls[c.g] : (real)0;

    // ---- backward sweep (K2 :283-366) fused with K3's per-class collection (:460-497)
    real Rb = NINF;
    for (int t = q.T - 1; t >= 0; --t) {
        if (c.active) {
            if (t == c.Tb - 1) {
                Rb = (c.s == 2 * c.L || (c.L > 0 && c.s == 2 * c.L - 1)) ? (real)0 : NINF;
            } else if (t < c.Tb - 1 && c.in_range) {
                const real *bq = As + ((t + 1) & 1) * rowStates + c.g * q.SS + c.s;
                Rb = lse3<FAST>(bq[0], c.s < 2 * c.L ? bq[1] : NINF, c.skip_bwd ? bq[2] : NINF);
            } else Rb = NINF;
            As[(t & 1) * rowStates + c.g * q.SS + c.s] = Rb + Qall[t * rowElems + c.g * q.C + c.cur];
            if (c.in_range && t < c.Tb) {
                const real v = Ra[t * rowStates + c.g * q.SS + c.s] + Rb;
                if (v != NINF) {
                    const int o = t * rowElems + c.g * q.C + c.cur;
                    pres[o] = 1;
                    atomicAdd(acc + o, ex<FAST>(v + my_nll));
                }
            }
        }
        __syncthreads();
    }

    // ---- factor: (1 - acc) [* go] where the class is present and t < Tb, else 0 (K3 :501-515)
    for (int i = tid; i < q.T * Gv * q.C; i += nth) {
        const int t = i / (Gv * q.C);
        const int r = i - t * (Gv * q.C);
        const int g = r / q.C;
        const int o = t * rowElems + r;
        real f = 0;
        if (pres[o] && t < il[b0 + g]) {
            f = (real)1 - acc[o];
            if (MODE == MODE_GRAD) f *= grad_out[(int64_t)(b0 + g) * go_stride];
        }
        if (MODE == MODE_FAC) fac_out[((int64_t)(b0 + g) * q.T + t) * q.C + (r - g * q.C)] = f;
        else acc[o] = f;
    }
    if (MODE == MODE_GRAD) {
        __syncthreads();
        if (q.vec == 4) phase_grad<real, FAST, (sizeof(real) == 4 ? 4 : 1)>(q, lp, acc, grad, b0, Gv);
        else if (q.vec == 2) phase_grad<real, FAST, (sizeof(real) == 8 ? 2 : 1)>(q, lp, acc, grad, b0, Gv);
        else phase_grad<real, FAST, 1>(q, lp, acc, grad, b0, Gv);
    }
}

// Training backward: grad[t,h,b,c] = exp(lp) * gfac[b,t,c] * go[b].  Pure streaming; thread = 16-byte vector.
template <bool FAST, int VE>
__global__ void ctc2d_apply_kernel(const float *__restrict__ lp, const float *__restrict__ fac,
                                   const float *__restrict__ go, int64_t go_stride, int T, int H, int N, int C,
                                   float *__restrict__ grad) {
    const int64_t row = (int64_t)N * C;           // one (t,h) row
    const int64_t nvec_row = row / VE;
    const int64_t total = (int64_t)T * H * nvec_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nvec_row;
        const int64_t j = (i - r * nvec_row) * VE;
        const int t = (int)(r / H);
        const int64_t off = r * row + j;
        VecT<float, VE> x = ld_cs<float, VE>(lp + off);
        VecT<float, VE> o;
#pragma unroll
        for (int k = 0; k < VE; ++k) {
            const int64_t e = j + k;
            const int b = (int)(e / C);
            const int cc = (int)(e - (int64_t)b * C);
            const float f = __ldg(fac + ((int64_t)b * T + t) * C + cc) * __ldg(go + (int64_t)b * go_stride);
            o.v[k] = (f == 0.f) ? 0.f : ex<FAST>(x.v[k]) * f;
        }
        st_cs<float, VE>(grad + off, o);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int check_common(const void *lp, const void *tg, const void *il, const void *tl, int64_t T, int64_t H, int64_t N,
                 int64_t C, int64_t S, int64_t blank) {
    if (T < 0 || H < 0 || N < 0 || C <= 0 || S < 0) return MR_ERR_BAD_SHAPE;
    if (blank < 0 || blank >= C) return MR_ERR_BLANK_RANGE;
    if (2 * S + 1 > 1024) return MR_ERR_TARGET_TOO_LONG;
    if (T > (1 << 24) || H > (1 << 20) || C > (1 << 24) || N > (1LL << 31) - 1) return MR_ERR_BAD_SHAPE;
    if (N > 0 && T > 0 && H > 0 && (!lp || !tg || !il || !tl)) return MR_ERR_NULL_POINTER;
    return MR_OK;
}

template <typename real>
int pick_vec(const void *p, int64_t N, int64_t C, int G) {
    const int ve = 16 / (int)sizeof(real);
    if (((uintptr_t)p % 16) == 0 && (N * C) % ve == 0 && ((int64_t)G * C) % ve == 0) return ve;
    return 1;
}

int smem_limit() {
    static int lim = -1;
    if (lim < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
            v = 48 * 1024;
        lim = v;
    }
    return lim;
}

template <typename real, bool FAST>
int launch_alpha(const real *lp, const int64_t *tg, const int64_t *il, const int64_t *tl, int64_t T, int64_t H,
                 int64_t N, int64_t C, int64_t S, int64_t tg_sn, int64_t tg_ss, int64_t blank, real *nll, real *la,
                 cudaStream_t st) {
    Geo q;
    q.T = (int)T; q.H = (int)H; q.N = (int)N; q.C = (int)C; q.S = (int)S; q.SS = (int)(2 * S + 1);
    q.blank = (int)blank; q.tg_sn = tg_sn; q.tg_ss = tg_ss;
    int G = 288 / q.SS;
    if (G < 1) G = 1;
    if (G > 8) G = 8;
    const size_t small = sizeof(real) * ((size_t)G * C + (size_t)G * q.SS + 2 * G);
    // staged path: kStages slabs of [H][G*C]; shrink G until it fits (<= ~56 KB keeps 4 CTAs/SM)
    bool staged = false;
    int Gs = G;
    for (; Gs >= 1; --Gs) {
        const size_t need = sizeof(real) * (size_t)kStages * H * Gs * C + sizeof(real) * ((size_t)Gs * C + (size_t)Gs * q.SS + 2 * Gs);
        if (need <= (size_t)56 * 1024 || (Gs == 1 && need <= (size_t)smem_limit())) { staged = true; break; }
    }
    size_t smem;
    if (staged) {
        G = Gs;
        smem = sizeof(real) * (size_t)kStages * H * G * C + sizeof(real) * ((size_t)G * C + (size_t)G * q.SS + 2 * G);
    } else {
        smem = small;
        if (smem > (size_t)smem_limit()) return MR_ERR_UNSUPPORTED;
    }
    q.G = G;
    q.vec = pick_vec<real>(lp, N, C, G);
    const int threads = (int)round_up((int64_t)G * q.SS, 32);
    const int grid = (int)ceil_div(N, G);
    auto kern = staged ? ctc2d_alpha_kernel<real, FAST, true> : ctc2d_alpha_kernel<real, FAST, false>;
    if (smem > 48 * 1024)
        MR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ctc2d_alpha attr");
    kern<<<grid, threads, smem, st>>>(q, lp, tg, il, tl, nll, la);
    return check_launch("ctc2d_alpha_kernel");
}

template <typename real, bool FAST, int MODE>
int launch_dp(const real *lp, const int64_t *tg, const int64_t *il, const int64_t *tl, const real *go,
              int64_t go_stride, int64_t T, int64_t H, int64_t N, int64_t C, int64_t S, int64_t tg_sn,
              int64_t tg_ss, int64_t blank, real *nll, real *fac, real *grad, cudaStream_t st) {
    Geo q;
    q.T = (int)T; q.H = (int)H; q.N = (int)N; q.C = (int)C; q.S = (int)S; q.SS = (int)(2 * S + 1);
    q.blank = (int)blank; q.tg_sn = tg_sn; q.tg_ss = tg_ss;
    int G = 160 / q.SS;  // fewer samples per CTA than the alpha kernel: the sweeps are latency-bound,
    if (G < 1) G = 1;    // more co-resident CTAs keep HBM busy meanwhile
    if (G > 8) G = 8;
    auto need = [&](int g) {
        return sizeof(real) * ((size_t)2 * T * g * C + (size_t)T * g * q.SS + (size_t)2 * g * q.SS + 3 * g) +
               (size_t)T * g * C + 16;
    };
    while (G > 1 && need(G) > (size_t)44 * 1024) --G;
    const size_t smem = need(G);
    if (smem > (size_t)smem_limit()) return MR_ERR_UNSUPPORTED;
    q.G = G;
    q.vec = pick_vec<real>(lp, N, C, G);
    if (MODE == MODE_GRAD && ((uintptr_t)grad % 16) != 0) q.vec = 1;
    const int threads = (int)round_up((int64_t)G * q.SS, 32);
    const int grid = (int)ceil_div(N, G);
    auto kern = ctc2d_dp_kernel<real, FAST, MODE>;
    if (smem > 48 * 1024)
        MR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ctc2d_dp attr");
    kern<<<grid, threads, smem, st>>>(q, lp, tg, il, tl, go, go_stride, nll, fac, grad);
    return check_launch("ctc2d_dp_kernel");
}

}  // namespace

extern "C" {

int mr_ctc2d_forward_f32(const float *lp, const int64_t *tg, const int64_t *il, const int64_t *tl, int64_t T,
                         int64_t H, int64_t N, int64_t C, int64_t S, int64_t tg_sn, int64_t tg_ss, int64_t blank,
                         int fast_math, float *nll, float *la, void *stream) {
    int rc = check_common(lp, tg, il, tl, T, H, N, C, S, blank);
    if (rc) return rc;
    if (N == 0) return MR_OK;
    if (T == 0 || H == 0) return MR_ERR_BAD_SHAPE;
    if (!nll || !la) return MR_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    return fast_math ? launch_alpha<float, true>(lp, tg, il, tl, T, H, N, C, S, tg_sn, tg_ss, blank, nll, la, st)
                     : launch_alpha<float, false>(lp, tg, il, tl, T, H, N, C, S, tg_sn, tg_ss, blank, nll, la, st);
}

int mr_ctc2d_forward_f64(const double *lp, const int64_t *tg, const int64_t *il, const int64_t *tl, int64_t T,
                         int64_t H, int64_t N, int64_t C, int64_t S, int64_t tg_sn, int64_t tg_ss, int64_t blank,
                         int fast_math, double *nll, double *la, void *stream) {
    (void)fast_math;
    int rc = check_common(lp, tg, il, tl, T, H, N, C, S, blank);
    if (rc) return rc;
    if (N == 0) return MR_OK;
    if (T == 0 || H == 0) return MR_ERR_BAD_SHAPE;
    if (!nll || !la) return MR_ERR_NULL_POINTER;
    return launch_alpha<double, false>(lp, tg, il, tl, T, H, N, C, S, tg_sn, tg_ss, blank, nll, la, (cudaStream_t)stream);
}

int mr_ctc2d_backward_f32(const float *go, int64_t go_stride, const float *lp, const int64_t *tg, const int64_t *il,
                          const int64_t *tl, const float *nll, const float *la, int64_t T, int64_t H, int64_t N,
                          int64_t C, int64_t S, int64_t tg_sn, int64_t tg_ss, int64_t blank, int fast_math,
                          float *grad, void *stream) {
    (void)nll; (void)la;
    int rc = check_common(lp, tg, il, tl, T, H, N, C, S, blank);
    if (rc) return rc;
    if (N == 0 || T == 0 || H == 0) return MR_OK;
    if (!go || !grad) return MR_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    return fast_math ? launch_dp<float, true, MODE_GRAD>(lp, tg, il, tl, go, go_stride, T, H, N, C, S, tg_sn, tg_ss, blank, nullptr, nullptr, grad, st)
                     : launch_dp<float, false, MODE_GRAD>(lp, tg, il, tl, go, go_stride, T, H, N, C, S, tg_sn, tg_ss, blank, nullptr, nullptr, grad, st);
}

int mr_ctc2d_backward_f64(const double *go, int64_t go_stride, const double *lp, const int64_t *tg, const int64_t *il,
                          const int64_t *tl, const double *nll, const double *la, int64_t T, int64_t H, int64_t N,
                          int64_t C, int64_t S, int64_t tg_sn, int64_t tg_ss, int64_t blank, int fast_math,
                          double *grad, void *stream) {
    (void)nll; (void)la; (void)fast_math;
    int rc = check_common(lp, tg, il, tl, T, H, N, C, S, blank);
    if (rc) return rc;
    if (N == 0 || T == 0 || H == 0) return MR_OK;
    if (!go || !grad) return MR_ERR_NULL_POINTER;
    return launch_dp<double, false, MODE_GRAD>(lp, tg, il, tl, go, go_stride, T, H, N, C, S, tg_sn, tg_ss, blank, nullptr, nullptr, grad, (cudaStream_t)stream);
}

int mr_ctc2d_forward_train_f32(const float *lp, const int64_t *tg, const int64_t *il, const int64_t *tl, int64_t T,
                               int64_t H, int64_t N, int64_t C, int64_t S, int64_t tg_sn, int64_t tg_ss,
                               int64_t blank, int fast_math, float *nll, float *gfac, void *stream) {
    int rc = check_common(lp, tg, il, tl, T, H, N, C, S, blank);
    if (rc) return rc;
    if (N == 0) return MR_OK;
    if (T == 0 || H == 0) return MR_ERR_BAD_SHAPE;
    if (!nll || !gfac) return MR_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    return fast_math ? launch_dp<float, true, MODE_FAC>(lp, tg, il, tl, nullptr, 0, T, H, N, C, S, tg_sn, tg_ss, blank, nll, gfac, nullptr, st)
                     : launch_dp<float, false, MODE_FAC>(lp, tg, il, tl, nullptr, 0, T, H, N, C, S, tg_sn, tg_ss, blank, nll, gfac, nullptr, st);
}

int mr_ctc2d_backward_apply_f32(const float *go, int64_t go_stride, const float *lp, const float *gfac, int64_t T,
                                int64_t H, int64_t N, int64_t C, int fast_math, float *grad, void *stream) {
    if (T < 0 || H < 0 || N < 0 || C <= 0) return MR_ERR_BAD_SHAPE;
    if (N == 0 || T == 0 || H == 0) return MR_OK;
    if (!go || !lp || !gfac || !grad) return MR_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = ((uintptr_t)lp % 16 == 0) && ((uintptr_t)grad % 16 == 0) && ((N * C) % 4 == 0);
    const int64_t nvec = T * H * (N * C / (v4 ? 4 : 1));
    const int threads = 256;
    int64_t blocks = ceil_div(nvec, threads);
    const int64_t cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    if (v4) {
        if (fast_math) ctc2d_apply_kernel<true, 4><<<(int)blocks, threads, 0, st>>>(lp, gfac, go, go_stride, (int)T, (int)H, (int)N, (int)C, grad);
        else ctc2d_apply_kernel<false, 4><<<(int)blocks, threads, 0, st>>>(lp, gfac, go, go_stride, (int)T, (int)H, (int)N, (int)C, grad);
    } else {
        if (fast_math) ctc2d_apply_kernel<true, 1><<<(int)blocks, threads, 0, st>>>(lp, gfac, go, go_stride, (int)T, (int)H, (int)N, (int)C, grad);
        else ctc2d_apply_kernel<false, 1><<<(int)blocks, threads, 0, st>>>(lp, gfac, go, go_stride, (int)T, (int)H, (int)N, (int)C, grad);
    }
    return check_launch("ctc2d_apply_kernel");
}

}  // extern "C"
