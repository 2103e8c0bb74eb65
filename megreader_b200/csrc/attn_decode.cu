// The recurrent loop of the attention recogniser head as persistent cooperative kernels.
//
//   reference: decoders/attention_decoder.py:96-131 (training loop with teacher forcing / step dropout, eval loop) around
//   AttentionRNNCell.forward (:187-231): per step ~15 framework kernels forward (embedding, Linear, cat + Linear + tanh + bmm for the
//   energies (:160-171), softmax, bmm for the context, GRUCell, Linear, (log-)softmax, NLL / argmax) and ~40 backward.
//
//   attn_fwd_kernel<false>   eval: greedy decoding, max_size steps, pred (+ per-step softmax)
//   attn_fwd_kernel<true>    training forward: fed-back symbol = target / own argmax / dropout noise from caller-made draws, masked
//                            NLL summed over the steps, attention maps; keeps per-step state for the backward
//   attn_bwd_kernel          training backward through time: gradients w.r.t. `projected`, `memory`, v, the word table; per-step
//                            pre-activation gradients for the weight-gradient GEMMs (which are plain dense GEMMs over S*N rows and
//                            are left to the caller)
//
// Forward, per step t (h_0 = 0, word_0 = blank):
//   P1  fh[n, :]    = Wa_h . h[n]                                    (the hidden half of the additive-attention Linear; the encoder
//                                                                      half `projected` is step-invariant and computed once by the caller)
//   P2  logits[n]   = Wout . h[n] + bout  -> output of the PREVIOUS step (eval: argmax -> pred[n, t-1]; training: log-softmax, NLL, and
//                     the symbol fed back)
//       score[l]    = v . tanh(projected[n, l] + fh[n]);  a = softmax_l(score);  context = sum_l a[l] memory[n, l]
//       x[n]        = [wordtab[word[n]] ; context]
//   P3  h'[n, j]    = GRU(x[n], h[n])_j                               (gi = W_ih x + b_ih, gh = W_hh h + b_hh, r, z, n gates)
// and one more P2 head after the last step.  P1 / P3 are parallel over output columns / hidden units: every CTA owns ceil(H / grid)
// of them and keeps ITS rows of Wa_h, W_ih, W_hh in shared memory for the whole loop (weights are read from HBM once per call, not
// once per step); a warp multiplies them with the activation rows of FOUR samples at a time (lane = k index; 4 units x 4 samples x
// 4 gate sums = 64 accumulators per lane, reduced with a 64-shuffle butterfly instead of 320 xor-shuffles), so that one shared-memory
// weight load feeds four FMAs.  The activation rows (written by the other CTAs in the previous phase) stream L2 -> shared memory
// through a warp-private 4-stage cp.async ring: the first version loaded them into registers inside the k loop and was bound by the
// bytes in flight (12 B/clk/SM at N = 256; one L2 round trip per iteration at N = 32).  With fewer sample groups than warps the
// reduction range is split over 2 or 4 warps and summed through shared memory.  P2 is parallel over samples.  The phases are separated by grid-wide barriers (monotonic arrival
// counter, cooperative launch so that co-residency is guaranteed; every wait is bounded and raises an error word instead of hanging).
// The backward runs the same structure in reverse with TRANSPOSED weight slices stationary in shared memory (see attn_bwd_kernel).
// All arithmetic is fp32 with accurate tanhf / expf / logf: the results match the framework composition to rounding, and the decoded
// strings are identical on the committed goldens.
#include "common.cuh"
#include <math.h>

namespace {
using namespace mr;

constexpr int kAttnThreads = 512;
constexpr int kTS = 4;            // samples per warp tile
constexpr int kUC = 4;            // weight rows (hidden units / output columns) per warp tile
constexpr unsigned kFull = 0xffffffffu;

struct AttnArgs {
    const float *projected;   // [N][L][H]   Wa_enc . memory + b   (bias included)
    const float *memory;      // [N][L][D]   D = H + E
    const float *wa_h;        // [H] rows of ld_wa floats: attn.attn.weight[:, :H]
    int64_t ld_wa;
    const float *v;           // [H]
    const float *wordtab;     // [V][H]      word_linear(embedding(v))
    const float *w_ih, *b_ih; // [3H][H + D], [3H]
    const float *w_hh, *b_hh; // [3H][H],     [3H]
    const float *w_out, *b_out;   // [V][H], [V]
    // state.  eval: h [2][N][H] ping-pong, fh [N][H], x [N][X].  training: one slice per step, h [S+1][N][H], fh [S][N][H], x [S][N][X]
    float *h, *fh, *x;
    int64_t h_stride;         // floats between two hidden-state slices
    // eval outputs
    int *pred;                // [N][S]
    float *prob;              // [N][S][V] softmax of the logits (the reference's per-step output), or nullptr
    // training inputs / outputs
    const int *targets;       // [N][S]
    const int *lengths;       // [N]
    const int *coin;          // [S]     1: feed the target back, 0: the step's own argmax
    const int *swap;          // [S][N]  1: replace the fed-back symbol by noise
    const int *noise;         // [S][N]
    float *gates;             // [S][N][4][H]  r, z, n, W_hn h + b_hn
    float *logp;              // [S][N][V]
    float *attn;              // [N][S][L]
    int *word;                // [S][N]  symbol fed into step t
    float *loss;              // [N], zeroed by the entry point
    unsigned *sync;           // [2]: arrival counter (zeroed), error word
    int N, L, H, D, V, S, blank, upc;   // upc = hidden units (and Wa_h columns) per CTA
    int Xp;                   // floats between two rows of x: 2H + E rounded up to a multiple of 4
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid-wide barrier number `idx` (0, 1, 2, ...): arrivals are counted monotonically
__device__ __forceinline__ void grid_barrier(unsigned *sync, unsigned idx) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(sync, 1u);
        const unsigned target = (idx + 1u) * gridDim.x;
        const long long t0 = clock64();
        while (ld_acquire_u32(sync) < target) {
            if (clock64() - t0 > (1ll << 31)) { atomicExch(sync + 1, 1u); break; }      // ~1 s: a peer is missing; flag it and go on
            __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(kFull, v, s);
    return v;
}

// Warp-wide sums of 16*Q per-lane partials in 16*Q - Q + Q shuffles: v[pair * Q + q], pair = 0..15.  Every stage folds the upper half
// of the index range onto the lanes whose bit is set and the lower half onto the others; afterwards lanes 2p and 2p+1 both hold the
// complete sums of pair p in v[0..Q).
template <int Q>
__device__ __forceinline__ void reduce_pairs(float (&v)[16 * Q], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8 * Q; ++i) {
            const float a = v[i], b = v[i + 8 * Q];
            v[i] = (up ? b : a) + __shfl_xor_sync(kFull, up ? a : b, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4 * Q; ++i) {
            const float a = v[i], b = v[i + 4 * Q];
            v[i] = (up ? b : a) + __shfl_xor_sync(kFull, up ? a : b, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2 * Q; ++i) {
            const float a = v[i], b = v[i + 2 * Q];
            v[i] = (up ? b : a) + __shfl_xor_sync(kFull, up ? a : b, 4);
        }
    }
    {
        const bool up = lane & 2;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const float a = v[i], b = v[i + Q];
            v[i] = (up ? b : a) + __shfl_xor_sync(kFull, up ? a : b, 2);
        }
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) v[i] += __shfl_xor_sync(kFull, v[i], 1);
}

// Streams kTS activation rows (global memory written by OTHER CTAs before the last grid barrier: rows base + min(n0 + s, N - 1) * ld,
// 16-byte aligned, readable up to round_up(k1, 4)) over k in [k0, k1) through a warp-private cp.async ring (L2 -> shared memory, no
// register staging: kStages - 1 chunks of kTS x kChunk floats are in flight per warp whatever the register budget) and calls
// body(k, xv[kTS]) for this lane's k values (k = k0 + lane, + 32, ...).  k0 must be a multiple of 4.
constexpr int kChunk = 64;        // floats of k per stage and row
constexpr int kStages = 4;
constexpr int kStageFloats = kStages * kTS * kChunk;     // per warp: 4 KB

template <typename Body>
__device__ __forceinline__ void staged_rows(float *stage, const float *base, int64_t ld, int n0, int N, int k0, int k1, int lane, Body &&body) {
    const int nch = (k1 - k0 + kChunk - 1) / kChunk;
    auto issue = [&](int c) {
        if (c < nch) {
            const int kc = k0 + c * kChunk;
            float *dst = stage + (c % kStages) * (kTS * kChunk);
#pragma unroll
            for (int j = 0; j < (kTS * kChunk / 4) / 32; ++j) {
                const int p = lane + 32 * j, row = p / (kChunk / 4), f4 = p % (kChunk / 4);
                const int k = kc + 4 * f4;
                if (k < k1) cp_async16(dst + row * kChunk + 4 * f4, base + (int64_t)min(n0 + row, N - 1) * ld + k);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < kStages - 1; ++c) issue(c);
    for (int c = 0; c < nch; ++c) {
        issue(c + kStages - 1);
        cp_async_wait<kStages - 1>();
        __syncwarp();
        const float *src = stage + (c % kStages) * (kTS * kChunk);
        const int kc = k0 + c * kChunk;
#pragma unroll
        for (int u = 0; u < kChunk / 32; ++u) {
            const int kk = lane + 32 * u, k = kc + kk;
            if (k < k1) {
                float xv[kTS];
#pragma unroll
                for (int s = 0; s < kTS; ++s) xv[s] = src[s * kChunk + kk];
                body(k, xv);
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

__device__ __forceinline__ float *align16(float *p) { return (float *)(((uintptr_t)p + 15) & ~(uintptr_t)15); }
constexpr size_t kRoundsSmemBytes = (kAttnThreads / 32) * 64 * sizeof(float) + 16 + (kAttnThreads / 32) * kStageFloats * sizeof(float);

// k-slices per sample group: with fewer groups than warps the reduction range is split over 2 or 4 warps
__device__ __forceinline__ int k_slices(int N) {
    const int ngroups = (N + kTS - 1) / kTS, nwarps = kAttnThreads / 32;
    return ngroups * 2 > nwarps ? 1 : (ngroups * 4 > nwarps ? 2 : 4);
}
__device__ __forceinline__ int slice_len(int K, int KS) { return ((K + KS - 1) / KS + kChunk - 1) / kChunk * kChunk; }

// One pass over all samples for one chunk of kUC weight rows: the work items (sample group g, k-slice ks) go round-robin to the
// warps; compute(g, ks, acc) accumulates the lane partials acc[(c * kTS + s) * Q + q]; the warp sums go to shared memory and
// finish(n, c, sums[Q]) runs once per (sample, weight row) on the total over the k-slices.  Called by all threads of the CTA.
template <int Q, typename Compute, typename Finish>
__device__ __forceinline__ void tile_rounds(float *s_part, int N, int KS, Compute &&compute, Finish &&finish) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kAttnThreads / 32;
    const int ngroups = (N + kTS - 1) / kTS, nitems = ngroups * KS;
    for (int base = 0; base < nitems; base += nwarps) {
        const int item = base + warp;
        if (item < nitems) {
            float acc[16 * Q];
#pragma unroll
            for (int i = 0; i < 16 * Q; ++i) acc[i] = 0.f;
            compute(item / KS, item % KS, acc);
            reduce_pairs<Q>(acc, lane);
            if (!(lane & 1)) {
#pragma unroll
                for (int q = 0; q < Q; ++q) s_part[(warp * 16 + (lane >> 1)) * Q + q] = acc[q];
            }
        }
        __syncthreads();
        const int g0 = base / KS, ng = min(nwarps / KS, ngroups - g0);
        for (int idx = threadIdx.x; idx < ng * 16; idx += kAttnThreads) {
            const int gl = idx >> 4, pr = idx & 15;
            float sums[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) sums[q] = 0.f;
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int q = 0; q < Q; ++q) sums[q] += s_part[((gl * KS + ks) * 16 + pr) * Q + q];
            }
            const int n = (g0 + gl) * kTS + pr % kTS;
            if (n < N) finish(n, pr / kTS, sums);
        }
        __syncthreads();
    }
}

template <bool TRAIN>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_fwd_kernel(AttnArgs a) {
    extern __shared__ float sm[];
    const int H = a.H, D = a.D, X = a.H + a.D, V = a.V, L = a.L, N = a.N, S = a.S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kAttnThreads / 32;
    // hidden units / Wa_h columns of this CTA
    const int j0 = min((int)blockIdx.x * a.upc, H), j1 = min(j0 + a.upc, H), nj = j1 - j0;
    // shared memory: [upc][H] Wa_h rows | [3 upc][X] W_ih rows | [3 upc][H] W_hh rows | scratch of phase 2
    float *s_wa = sm;
    float *s_wih = s_wa + (size_t)a.upc * H;
    float *s_whh = s_wih + (size_t)3 * a.upc * X;
    float *s_p2 = s_whh + (size_t)3 * a.upc * H;       // fh / h row (H), scores (L), logits (V)
    float *s_part = s_p2 + H + L + V;                   // [warps][16 pairs][4] warp sums of a round
    float *s_stage = align16(s_part + (kAttnThreads / 32) * 64) + (size_t)warp * kStageFloats;   // this warp's cp.async ring
    for (int i = threadIdx.x; i < nj * H; i += kAttnThreads) s_wa[i] = a.wa_h[(int64_t)(j0 + i / H) * a.ld_wa + i % H];
    for (int g = 0; g < 3; ++g) {
        for (int i = threadIdx.x; i < nj * X; i += kAttnThreads) s_wih[(size_t)(g * a.upc) * X + i] = a.w_ih[((int64_t)g * H + j0) * X + i];
        for (int i = threadIdx.x; i < nj * H; i += kAttnThreads) s_whh[(size_t)(g * a.upc) * H + i] = a.w_hh[((int64_t)g * H + j0) * H + i];
    }
    __syncthreads();
    float *s_row = s_p2, *s_score = s_p2 + H, *s_logit = s_score + L;
    __shared__ int s_word;
    unsigned bar = 0;
    const int Xp = a.Xp;                                  // row stride of x (X rounded up to 4 floats: 16-byte rows for cp.async)
    const int64_t NH = (int64_t)N * H, NX = (int64_t)N * Xp;
    const int KS = k_slices(N), slx = slice_len(X, KS), slh = slice_len(H, KS);

    for (int t = 0; t <= S; ++t) {
        const float *h = a.h + (TRAIN ? t : (t & 1)) * a.h_stride;     // hidden state after t steps
        float *hn = a.h + (TRAIN ? t + 1 : ((t + 1) & 1)) * a.h_stride;
        float *fh = a.fh + (TRAIN ? t : 0) * NH;
        float *x = a.x + (TRAIN ? t : 0) * NX;
        // ---------------- P1: fh = Wa_h . h (columns j0..j1 of every sample); skipped after the last step
        if (t < S) {
            for (int cb = 0; cb < nj; cb += kUC) {
                const float *wr[kUC];
#pragma unroll
                for (int c = 0; c < kUC; ++c) wr[c] = s_wa + (size_t)min(cb + c, nj - 1) * H;
                tile_rounds<1>(s_part, N, KS,
                    [&](int g, int ks, float (&acc)[16]) {
                        staged_rows(s_stage, h, H, g * kTS, N, ks * slh, min(H, (ks + 1) * slh), lane, [&](int k, const float (&xv)[kTS]) {
#pragma unroll
                            for (int c = 0; c < kUC; ++c) {
                                const float w = wr[c][k];
#pragma unroll
                                for (int s = 0; s < kTS; ++s) acc[c * kTS + s] = fmaf(w, xv[s], acc[c * kTS + s]);
                            }
                        });
                    },
                    [&](int n, int c, const float (&sums)[1]) {
                        if (cb + c < nj) fh[(int64_t)n * H + j0 + cb + c] = sums[0];
                    });
            }
            grid_barrier(a.sync, bar++);
        }
        // ---------------- P2: per sample -- previous step's output symbol, attention, context, GRU input
        for (int n = blockIdx.x; n < N; n += gridDim.x) {
            __syncthreads();
            for (int i = threadIdx.x; i < H; i += kAttnThreads) s_row[i] = __ldcg(h + (int64_t)n * H + i);
            __syncthreads();
            if (t > 0) {
                for (int vv = warp; vv < V; vv += nwarps) {
                    float acc = 0.f;
                    for (int k = lane; k < H; k += 32) acc = fmaf(__ldg(a.w_out + (int64_t)vv * H + k), s_row[k], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) s_logit[vv] = acc + a.b_out[vv];
                }
                __syncthreads();
                if (warp == 0) {
                    float best = -INFINITY;
                    int bi = 0x7fffffff;
                    for (int vv = lane; vv < V; vv += 32)
                        if (s_logit[vv] > best) { best = s_logit[vv]; bi = vv; }
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) {
                        const float ob = __shfl_xor_sync(kFull, best, s);
                        const int oi = __shfl_xor_sync(kFull, bi, s);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }       // first maximum, like torch.argmax
                    }
                    float den = 0.f;
                    for (int vv = lane; vv < V; vv += 32) den += expf(s_logit[vv] - best);
                    den = warp_sum(den);
                    if (TRAIN) {
                        const float lden = logf(den);
                        float *lp = a.logp + ((int64_t)(t - 1) * N + n) * V;
                        for (int vv = lane; vv < V; vv += 32) lp[vv] = (s_logit[vv] - best) - lden;
                        if (lane == 0) {
                            const int tgt = min(max(a.targets[(int64_t)n * S + t - 1], 0), V - 1);
                            if (t - 1 <= a.lengths[n]) a.loss[n] -= (s_logit[tgt] - best) - lden;      // NLLLoss * (t <= lengths)
                            int w = a.coin[t - 1] ? tgt : bi;
                            if (a.swap[(int64_t)(t - 1) * N + n]) w = min(max(a.noise[(int64_t)(t - 1) * N + n], 0), V - 1);
                            s_word = w;
                        }
                    } else {
                        if (lane == 0) { s_word = bi; a.pred[(int64_t)n * S + t - 1] = bi; }
                        if (a.prob)
                            for (int vv = lane; vv < V; vv += 32) a.prob[((int64_t)n * S + t - 1) * V + vv] = expf(s_logit[vv] - best) / den;
                    }
                }
            } else if (threadIdx.x == 0) {
                s_word = a.blank;
            }
            if (t == S) continue;
            __syncthreads();
            if (TRAIN && threadIdx.x == 0) a.word[(int64_t)t * N + n] = s_word;
            for (int i = threadIdx.x; i < H; i += kAttnThreads) s_row[i] = __ldcg(fh + (int64_t)n * H + i);
            __syncthreads();
            const float *pj = a.projected + (int64_t)n * L * H;
            for (int l = warp; l < L; l += nwarps) {
                float acc = 0.f;
                for (int k = lane; k < H; k += 32) acc = fmaf(a.v[k], tanhf(pj[(int64_t)l * H + k] + s_row[k]), acc);
                acc = warp_sum(acc);
                if (lane == 0) s_score[l] = acc;
            }
            __syncthreads();
            if (warp == 0) {
                float mx = -INFINITY;
                for (int l = lane; l < L; l += 32) mx = fmaxf(mx, s_score[l]);
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, s));
                float den = 0.f;
                for (int l = lane; l < L; l += 32) den += expf(s_score[l] - mx);
                den = warp_sum(den);
                for (int l = lane; l < L; l += 32) {
                    const float w = expf(s_score[l] - mx) / den;
                    s_score[l] = w;
                    if (TRAIN) a.attn[((int64_t)n * S + t) * L + l] = w;
                }
            }
            __syncthreads();
            const float *mem = a.memory + (int64_t)n * L * D;
            float *xr = x + (int64_t)n * Xp;
            for (int d = threadIdx.x; d < D; d += kAttnThreads) {
                float acc = 0.f;
                for (int l = 0; l < L; ++l) acc = fmaf(s_score[l], mem[(int64_t)l * D + d], acc);
                xr[H + d] = acc;
            }
            const float *wt = a.wordtab + (int64_t)s_word * H;
            for (int i = threadIdx.x; i < H; i += kAttnThreads) xr[i] = wt[i];
        }
        if (t == S) break;
        grid_barrier(a.sync, bar++);
        // ---------------- P3: GRU cell for hidden units j0..j1 of every sample (torch.nn.GRUCell gate order r, z, n)
        for (int cb = 0; cb < nj; cb += kUC) {
            // per (unit c, sample s): q = 0 r-gate sum (input + hidden), 1 z-gate sum, 2 n-gate input part, 3 n-gate hidden part
            const float *wi[kUC], *wh[kUC];
#pragma unroll
            for (int c = 0; c < kUC; ++c) {
                wi[c] = s_wih + (size_t)min(cb + c, nj - 1) * X;
                wh[c] = s_whh + (size_t)min(cb + c, nj - 1) * H;
            }
            const size_t gx = (size_t)a.upc * X, gh = (size_t)a.upc * H;      // gate stride inside the shared-memory slices
            tile_rounds<4>(s_part, N, KS,
                [&](int g, int ks, float (&acc)[64]) {
                    staged_rows(s_stage, x, Xp, g * kTS, N, ks * slx, min(X, (ks + 1) * slx), lane, [&](int k, const float (&xv)[kTS]) {
#pragma unroll
                        for (int c = 0; c < kUC; ++c) {
                            const float w0 = wi[c][k], w1 = wi[c][gx + k], w2 = wi[c][2 * gx + k];
#pragma unroll
                            for (int s = 0; s < kTS; ++s) {
                                float *q = acc + (c * kTS + s) * 4;
                                q[0] = fmaf(w0, xv[s], q[0]);
                                q[1] = fmaf(w1, xv[s], q[1]);
                                q[2] = fmaf(w2, xv[s], q[2]);
                            }
                        }
                    });
                    staged_rows(s_stage, h, H, g * kTS, N, ks * slh, min(H, (ks + 1) * slh), lane, [&](int k, const float (&hv)[kTS]) {
#pragma unroll
                        for (int c = 0; c < kUC; ++c) {
                            const float w0 = wh[c][k], w1 = wh[c][gh + k], w2 = wh[c][2 * gh + k];
#pragma unroll
                            for (int s = 0; s < kTS; ++s) {
                                float *q = acc + (c * kTS + s) * 4;
                                q[0] = fmaf(w0, hv[s], q[0]);
                                q[1] = fmaf(w1, hv[s], q[1]);
                                q[3] = fmaf(w2, hv[s], q[3]);
                            }
                        }
                    });
                },
                [&](int n, int c, const float (&sums)[4]) {
                    if (cb + c >= nj) return;
                    const int j = j0 + cb + c;
                    const float r = 1.f / (1.f + expf(-(sums[0] + a.b_ih[j] + a.b_hh[j])));
                    const float z = 1.f / (1.f + expf(-(sums[1] + a.b_ih[H + j] + a.b_hh[H + j])));
                    const float ghn = sums[3] + a.b_hh[2 * H + j];
                    const float nn = tanhf(sums[2] + a.b_ih[2 * H + j] + r * ghn);
                    hn[(int64_t)n * H + j] = (1.f - z) * nn + z * __ldcg(h + (int64_t)n * H + j);
                    if (TRAIN) {
                        float *gt = a.gates + ((int64_t)t * N + n) * 4 * H + j;
                        gt[0] = r; gt[H] = z; gt[2 * H] = nn; gt[3 * H] = ghn;
                    }
                });
        }
        grid_barrier(a.sync, bar++);
    }
}

// ------------------------------------------------------------------------------------------------------------------ backward
// Reverse loop over the steps with dh = d loss / d h_t carried in global memory.  Per step:
//   B1 (per sample)   dlogits = (softmax - onehot(target)) * mask * grad_loss;  dh += dlogits . Wout;  GRU cell derivative ->
//                     dgi, dgh (pre-activation gradients of the two gate products), dh <- dh * z (direct path to h_{t-1})
//   B2 (per column)   dx = dgi . W_ih (GRU input gradient: word part + context part), dh += dgh . W_hh        [transposed slices]
//   B3 (per sample)   context / softmax / energy derivative: dmemory += a (x) dctx, dprojected += dpre, dv += ds . e, dfh = sum_l dpre;
//                     word-table rows += dx[:H]
//   B4 (per column)   dh += dfh . Wa_h                                                                           [transposed slice]
// Every CTA owns cx columns of W_ih (as rows of W_ih^T), ch columns of W_hh and of Wa_h, stationary in shared memory.
struct AttnBwdArgs {
    const float *projected, *memory, *wa_h;
    int64_t ld_wa;
    const float *v, *w_ih, *w_hh, *w_out;
    const float *h, *fh, *gates, *logp, *attn;     // saved by the forward
    const int *word, *targets, *lengths;
    const float *gloss;       // [N] upstream gradient of the per-sample loss
    float *dlogits;           // [S][N][V]
    float *dgi, *dgh;         // [S][N][3H]
    float *dfh;               // [S][N][H]
    float *dx;                // [N][X] scratch
    float *dh;                // [N][H], zeroed by the entry point
    float *dP, *dM;           // [N][L][H], [N][L][D], zeroed
    float *dv;                // [H], zeroed
    float *dwordtab;          // [V][H], zeroed
    unsigned *sync;
    int N, L, H, D, V, S, cx, ch;
};

__global__ void __launch_bounds__(kAttnThreads, 1) attn_bwd_kernel(AttnBwdArgs a) {
    extern __shared__ float sm[];
    const int H = a.H, D = a.D, X = a.H + a.D, V = a.V, L = a.L, N = a.N, S = a.S, H3 = 3 * a.H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kAttnThreads / 32;
    const int kx0 = min((int)blockIdx.x * a.cx, X), ncx = min(kx0 + a.cx, X) - kx0;       // columns of W_ih
    const int kh0 = min((int)blockIdx.x * a.ch, H), nch = min(kh0 + a.ch, H) - kh0;       // columns of W_hh / Wa_h
    float *s_wihT = sm;                                   // [cx][3H]
    float *s_whhT = s_wihT + (size_t)a.cx * H3;           // [ch][3H]
    float *s_waT = s_whhT + (size_t)a.ch * H3;            // [ch][H]
    float *s_dv = s_waT + (size_t)a.ch * H;               // [H]
    float *s_dctx = s_dv + H;                             // [D]
    float *s_a = s_dctx + D, *s_da = s_a + L, *s_ds = s_da + L;   // [L] each
    float *s_dlogit = s_ds + L;                           // [V]
    float *s_part = s_dlogit + V;                         // [warps][16 pairs] warp sums of a round
    float *s_stage = align16(s_part + (kAttnThreads / 32) * 64) + (size_t)warp * kStageFloats;   // this warp's cp.async ring
    for (int i = threadIdx.x; i < ncx * H3; i += kAttnThreads) s_wihT[i] = a.w_ih[(int64_t)(i % H3) * X + kx0 + i / H3];
    for (int i = threadIdx.x; i < nch * H3; i += kAttnThreads) s_whhT[i] = a.w_hh[(int64_t)(i % H3) * H + kh0 + i / H3];
    for (int i = threadIdx.x; i < nch * H; i += kAttnThreads) s_waT[i] = a.wa_h[(int64_t)(i % H) * a.ld_wa + kh0 + i / H];
    for (int i = threadIdx.x; i < H; i += kAttnThreads) s_dv[i] = 0.f;
    __syncthreads();
    unsigned bar = 0;
    const int KS = k_slices(N), sl3 = slice_len(H3, KS), slh = slice_len(H, KS);

    for (int t = S - 1; t >= 0; --t) {
        const int64_t tN = (int64_t)t * N;
        // ---------------- B1: output layer + GRU cell derivative, per sample
        for (int n = blockIdx.x; n < N; n += gridDim.x) {
            __syncthreads();
            const int tgt = min(max(a.targets[(int64_t)n * S + t], 0), V - 1);
            const float m = (t <= a.lengths[n]) ? a.gloss[n] : 0.f;
            for (int vv = threadIdx.x; vv < V; vv += kAttnThreads) {
                const float g = (expf(a.logp[(tN + n) * V + vv]) - (vv == tgt ? 1.f : 0.f)) * m;
                s_dlogit[vv] = g;
                a.dlogits[(tN + n) * V + vv] = g;
            }
            __syncthreads();
            const float *gt = a.gates + (tN + n) * 4 * H;
            const float *hp = a.h + (tN + n) * H;                     // h_{t-1} = slice t
            for (int k = threadIdx.x; k < H; k += kAttnThreads) {
                float d = __ldcg(a.dh + (int64_t)n * H + k);
                for (int vv = 0; vv < V; ++vv) d = fmaf(s_dlogit[vv], __ldg(a.w_out + (int64_t)vv * H + k), d);
                const float r = gt[k], z = gt[H + k], nn = gt[2 * H + k], ghn = gt[3 * H + k];
                const float dn_pre = d * (1.f - z) * (1.f - nn * nn);
                const float dz_pre = d * (hp[k] - nn) * z * (1.f - z);
                const float dr_pre = dn_pre * ghn * r * (1.f - r);
                float *gi = a.dgi + (tN + n) * H3, *gh = a.dgh + (tN + n) * H3;
                gi[k] = dr_pre; gi[H + k] = dz_pre; gi[2 * H + k] = dn_pre;
                gh[k] = dr_pre; gh[H + k] = dz_pre; gh[2 * H + k] = dn_pre * r;
                a.dh[(int64_t)n * H + k] = d * z;
            }
        }
        grid_barrier(a.sync, bar++);
        // ---------------- B2: dx = dgi . W_ih (columns kx0..), dh += dgh . W_hh (columns kh0..)
        for (int pass = 0; pass < 2; ++pass) {               // 0: dx from dgi and W_ih^T, 1: dh from dgh and W_hh^T
            const int ncol = pass ? nch : ncx;
            const float *sw = pass ? s_whhT : s_wihT;
            const float *act = (pass ? a.dgh : a.dgi) + tN * H3;
            for (int cb = 0; cb < ncol; cb += kUC) {
                const float *wr[kUC];
#pragma unroll
                for (int c = 0; c < kUC; ++c) wr[c] = sw + (size_t)min(cb + c, ncol - 1) * H3;
                tile_rounds<1>(s_part, N, KS,
                    [&](int g, int ks, float (&acc)[16]) {
                        staged_rows(s_stage, act, H3, g * kTS, N, ks * sl3, min(H3, (ks + 1) * sl3), lane, [&](int k, const float (&xv)[kTS]) {
#pragma unroll
                            for (int c = 0; c < kUC; ++c) {
                                const float w = wr[c][k];
#pragma unroll
                                for (int s = 0; s < kTS; ++s) acc[c * kTS + s] = fmaf(w, xv[s], acc[c * kTS + s]);
                            }
                        });
                    },
                    [&](int n, int c, const float (&sums)[1]) {
                        if (cb + c >= ncol) return;
                        if (pass == 0) {
                            a.dx[(int64_t)n * X + kx0 + cb + c] = sums[0];
                        } else {
                            float *p = a.dh + (int64_t)n * H + kh0 + cb + c;
                            *p = __ldcg(p) + sums[0];
                        }
                    });
            }
        }
        grid_barrier(a.sync, bar++);
        // ---------------- B3: attention derivative, per sample
        for (int n = blockIdx.x; n < N; n += gridDim.x) {
            __syncthreads();
            for (int d = threadIdx.x; d < D; d += kAttnThreads) s_dctx[d] = __ldcg(a.dx + (int64_t)n * X + H + d);
            for (int l = threadIdx.x; l < L; l += kAttnThreads) s_a[l] = a.attn[((int64_t)n * S + t) * L + l];
            {
                float *row = a.dwordtab + (int64_t)a.word[tN + n] * H;
                for (int k = threadIdx.x; k < H; k += kAttnThreads) atomicAdd(row + k, __ldcg(a.dx + (int64_t)n * X + k));
            }
            __syncthreads();
            const float *mem = a.memory + (int64_t)n * L * D;
            for (int l = warp; l < L; l += nwarps) {
                float acc = 0.f;
                for (int d = lane; d < D; d += 32) acc = fmaf(s_dctx[d], mem[(int64_t)l * D + d], acc);
                acc = warp_sum(acc);
                if (lane == 0) s_da[l] = acc;
            }
            __syncthreads();
            for (int l = threadIdx.x; l < L; l += kAttnThreads) {
                float dot = 0.f;
                for (int i = 0; i < L; ++i) dot = fmaf(s_a[i], s_da[i], dot);
                s_ds[l] = s_a[l] * (s_da[l] - dot);
            }
            float *dm = a.dM + (int64_t)n * L * D;
            for (int i = threadIdx.x; i < L * D; i += kAttnThreads) dm[i] += s_a[i / D] * s_dctx[i % D];
            __syncthreads();
            const float *pj = a.projected + (int64_t)n * L * H;
            float *dp = a.dP + (int64_t)n * L * H;
            for (int k = threadIdx.x; k < H; k += kAttnThreads) {
                const float fhk = a.fh[(tN + n) * H + k], vk = a.v[k];
                float dvk = 0.f, dfhk = 0.f;
                for (int l0 = 0; l0 < L; l0 += 8) {               // eight rows at a time: all loads issued before the first store
                    float pv[8], dq[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int l = min(l0 + u, L - 1);
                        pv[u] = __ldg(pj + (int64_t)l * H + k);
                        dq[u] = dp[(int64_t)l * H + k];
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (l0 + u < L) {
                            const float e = tanhf(pv[u] + fhk);
                            dvk = fmaf(s_ds[l0 + u], e, dvk);
                            const float dpre = s_ds[l0 + u] * vk * (1.f - e * e);
                            dp[(int64_t)(l0 + u) * H + k] = dq[u] + dpre;
                            dfhk += dpre;
                        }
                    }
                }
                a.dfh[(tN + n) * H + k] = dfhk;
                s_dv[k] += dvk;
            }
        }
        grid_barrier(a.sync, bar++);
        // ---------------- B4: dh += dfh . Wa_h (columns kh0..)
        for (int cb = 0; cb < nch; cb += kUC) {
            const float *wr[kUC];
#pragma unroll
            for (int c = 0; c < kUC; ++c) wr[c] = s_waT + (size_t)min(cb + c, nch - 1) * H;
            const float *act = a.dfh + tN * H;
            tile_rounds<1>(s_part, N, KS,
                [&](int g, int ks, float (&acc)[16]) {
                    staged_rows(s_stage, act, H, g * kTS, N, ks * slh, min(H, (ks + 1) * slh), lane, [&](int k, const float (&xv)[kTS]) {
#pragma unroll
                        for (int c = 0; c < kUC; ++c) {
                            const float w = wr[c][k];
#pragma unroll
                            for (int s = 0; s < kTS; ++s) acc[c * kTS + s] = fmaf(w, xv[s], acc[c * kTS + s]);
                        }
                    });
                },
                [&](int n, int c, const float (&sums)[1]) {
                    if (cb + c >= nch) return;
                    float *p = a.dh + (int64_t)n * H + kh0 + cb + c;
                    *p = __ldcg(p) + sums[0];
                });
        }
        grid_barrier(a.sync, bar++);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < H; k += kAttnThreads)
        if (s_dv[k] != 0.f) atomicAdd(a.dv + k, s_dv[k]);
}

template <typename... Args>
cudaError_t launch_coop(void (*kern)(Args...), dim3 grid, int threads, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)threads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

constexpr size_t kAttnSmemMax = 220 * 1024;

template <bool TRAIN>
int launch_fwd(AttnArgs &a, cudaStream_t st) {
    const int X = a.H + a.D;
    const int grid = sm_count();
    a.upc = (int)ceil_div(a.H, grid);
    const size_t smem = ((size_t)a.upc * a.H + (size_t)3 * a.upc * X + (size_t)3 * a.upc * a.H + a.H + a.L + a.V) * sizeof(float) + kRoundsSmemBytes;
    if (a.H % 4) return MR_ERR_UNSUPPORTED;               // 16-byte activation rows
    if (smem > kAttnSmemMax) return MR_ERR_UNSUPPORTED;
    int rc = ensure_dyn_smem((const void *)attn_fwd_kernel<TRAIN>, smem, "attn_fwd smem attr");
    if (rc) return rc;
    int max_blocks = 0;
    MR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, attn_fwd_kernel<TRAIN>, kAttnThreads, smem), "attn_fwd occupancy");
    if (max_blocks < 1) return MR_ERR_UNSUPPORTED;
    MR_CUDA_TRY(launch_coop(attn_fwd_kernel<TRAIN>, dim3((unsigned)grid), kAttnThreads, smem, st, a), "attn_fwd_kernel");
    return check_launch("attn_fwd_kernel");
}

}  // namespace

extern "C" {

/* floats of scratch the decode loop needs: h (2 x N x H), fh (N x H), x (N x (2H + E)), word (N ints), sync (2 words, 256 B) */
int64_t mr_attn_decode_workspace_bytes(int64_t N, int64_t H, int64_t E) {
    return round_up(N * H * 4, 256) * 3 + round_up(N * round_up(2 * H + E, 4) * 4, 256) + round_up(N * 4, 256) + 256;
}

/* Greedy decoding loop of AttentionDecoder.forward (eval branch, decoders/attention_decoder.py:119-131).  See the header. */
int mr_attn_decode_f32(const float *projected, const float *memory, const float *wa_h, int64_t ld_wa, const float *v,
                       const float *wordtab, const float *w_ih, const float *b_ih, const float *w_hh, const float *b_hh,
                       const float *w_out, const float *b_out, int *pred, float *prob, void *workspace, int64_t workspace_bytes,
                       int N, int L, int H, int E, int V, int S, int blank, void *stream) {
    if (N < 0 || L <= 0 || H <= 0 || E < 0 || V <= 0 || S <= 0 || blank < 0 || blank >= V) return MR_ERR_BAD_SHAPE;
    if (N == 0) return MR_OK;
    if (!projected || !memory || !wa_h || !v || !wordtab || !w_ih || !b_ih || !w_hh || !b_hh || !w_out || !b_out || !pred || !workspace)
        return MR_ERR_NULL_POINTER;
    if (workspace_bytes < mr_attn_decode_workspace_bytes(N, H, E) || ((uintptr_t)workspace % 256)) return MR_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    AttnArgs a = {};
    a.projected = projected; a.memory = memory; a.wa_h = wa_h; a.ld_wa = ld_wa; a.v = v; a.wordtab = wordtab;
    a.w_ih = w_ih; a.b_ih = b_ih; a.w_hh = w_hh; a.b_hh = b_hh; a.w_out = w_out; a.b_out = b_out;
    a.N = N; a.L = L; a.H = H; a.D = H + E; a.V = V; a.S = S; a.blank = blank; a.pred = pred; a.prob = prob;
    unsigned char *ws = (unsigned char *)workspace;
    const int64_t hb = round_up((int64_t)N * H * 4, 256), xb = round_up((int64_t)N * round_up(2 * H + E, 4) * 4, 256), wb = round_up((int64_t)N * 4, 256);
    a.Xp = (int)round_up(2 * H + E, 4);
    a.h = (float *)ws; a.h_stride = hb / 4; a.fh = (float *)(ws + 2 * hb); a.x = (float *)(ws + 3 * hb);
    a.sync = (unsigned *)(ws + 3 * hb + xb + wb);
    MR_CUDA_TRY(cudaMemsetAsync(a.h, 0, (size_t)N * H * 4, st), "cudaMemsetAsync(attn h0)");
    MR_CUDA_TRY(cudaMemsetAsync(a.sync, 0, 256, st), "cudaMemsetAsync(attn sync)");
    return launch_fwd<false>(a, st);
}

/* error word of the last decode on this workspace (0 = fine, 1 = a grid barrier timed out); synchronises the stream */
int mr_attn_decode_status(const void *workspace, int64_t N, int64_t H, int64_t E, void *stream, int *status) {
    if (!workspace || !status) return MR_ERR_NULL_POINTER;
    const unsigned char *ws = (const unsigned char *)workspace;
    const int64_t hb = round_up(N * H * 4, 256), xb = round_up(N * round_up(2 * H + E, 4) * 4, 256), wb = round_up(N * 4, 256);
    unsigned words[2] = {0, 0};
    MR_CUDA_TRY(cudaMemcpyAsync(words, ws + 3 * hb + xb + wb, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync(attn status)");
    MR_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize(attn status)");
    *status = (int)words[1];
    return MR_OK;
}

/* Training forward of the attention head's loop (decoders/attention_decoder.py:96-117).  See include/megreader_b200.h. */
int mr_attn_train_fwd_f32(const float *projected, const float *memory, const float *wa_h, int64_t ld_wa, const float *v,
                          const float *wordtab, const float *w_ih, const float *b_ih, const float *w_hh, const float *b_hh,
                          const float *w_out, const float *b_out, const int *targets, const int *lengths, const int *coin,
                          const int *swap, const int *noise, float *h_all, float *fh_all, float *x_all, float *gates, float *logp,
                          float *attn, int *word, float *loss, void *sync, int N, int L, int H, int E, int V, int S, int blank,
                          void *stream) {
    if (N < 0 || L <= 0 || H <= 0 || E < 0 || V <= 0 || S <= 0 || blank < 0 || blank >= V) return MR_ERR_BAD_SHAPE;
    if (N == 0) return MR_OK;
    if (!projected || !memory || !wa_h || !v || !wordtab || !w_ih || !b_ih || !w_hh || !b_hh || !w_out || !b_out || !targets || !lengths ||
        !coin || !swap || !noise || !h_all || !fh_all || !x_all || !gates || !logp || !attn || !word || !loss || !sync)
        return MR_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    AttnArgs a = {};
    a.projected = projected; a.memory = memory; a.wa_h = wa_h; a.ld_wa = ld_wa; a.v = v; a.wordtab = wordtab;
    a.w_ih = w_ih; a.b_ih = b_ih; a.w_hh = w_hh; a.b_hh = b_hh; a.w_out = w_out; a.b_out = b_out;
    a.N = N; a.L = L; a.H = H; a.D = H + E; a.V = V; a.S = S; a.blank = blank;
    a.targets = targets; a.lengths = lengths; a.coin = coin; a.swap = swap; a.noise = noise;
    a.Xp = (int)round_up(2 * H + E, 4);
    a.h = h_all; a.h_stride = (int64_t)N * H; a.fh = fh_all; a.x = x_all; a.gates = gates; a.logp = logp; a.attn = attn; a.word = word; a.loss = loss;
    a.sync = (unsigned *)sync;
    MR_CUDA_TRY(cudaMemsetAsync(h_all, 0, (size_t)N * H * 4, st), "cudaMemsetAsync(attn h0)");
    MR_CUDA_TRY(cudaMemsetAsync(loss, 0, (size_t)N * 4, st), "cudaMemsetAsync(attn loss)");
    MR_CUDA_TRY(cudaMemsetAsync(sync, 0, 8, st), "cudaMemsetAsync(attn sync)");
    return launch_fwd<true>(a, st);
}

/* Training backward of the attention head's loop.  See include/megreader_b200.h. */
int mr_attn_train_bwd_f32(const float *projected, const float *memory, const float *wa_h, int64_t ld_wa, const float *v,
                          const float *w_ih, const float *w_hh, const float *w_out, const float *h_all, const float *fh_all,
                          const float *gates, const float *logp, const float *attn, const int *word, const int *targets,
                          const int *lengths, const float *grad_loss, float *dlogits, float *dgi, float *dgh, float *dfh, float *dx,
                          float *dh, float *dprojected, float *dmemory, float *dv, float *dwordtab, void *sync, int N, int L, int H,
                          int E, int V, int S, void *stream) {
    if (N < 0 || L <= 0 || H <= 0 || E < 0 || V <= 0 || S <= 0) return MR_ERR_BAD_SHAPE;
    if (N == 0) return MR_OK;
    if (!projected || !memory || !wa_h || !v || !w_ih || !w_hh || !w_out || !h_all || !fh_all || !gates || !logp || !attn || !word ||
        !targets || !lengths || !grad_loss || !dlogits || !dgi || !dgh || !dfh || !dx || !dh || !dprojected || !dmemory || !dv || !dwordtab ||
        !sync)
        return MR_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    AttnBwdArgs a = {};
    a.projected = projected; a.memory = memory; a.wa_h = wa_h; a.ld_wa = ld_wa; a.v = v; a.w_ih = w_ih; a.w_hh = w_hh; a.w_out = w_out;
    a.h = h_all; a.fh = fh_all; a.gates = gates; a.logp = logp; a.attn = attn; a.word = word; a.targets = targets; a.lengths = lengths;
    a.gloss = grad_loss; a.dlogits = dlogits; a.dgi = dgi; a.dgh = dgh; a.dfh = dfh; a.dx = dx; a.dh = dh; a.dP = dprojected;
    a.dM = dmemory; a.dv = dv; a.dwordtab = dwordtab; a.sync = (unsigned *)sync;
    a.N = N; a.L = L; a.H = H; a.D = H + E; a.V = V; a.S = S;
    const int X = 2 * H + E, D = H + E;
    const int grid = sm_count();
    a.cx = (int)ceil_div(X, grid);
    a.ch = (int)ceil_div(H, grid);
    const size_t smem = ((size_t)a.cx * 3 * H + (size_t)a.ch * 3 * H + (size_t)a.ch * H + H + D + 3 * (size_t)L + V) * sizeof(float) + kRoundsSmemBytes;
    if (H % 4) return MR_ERR_UNSUPPORTED;
    if (smem > kAttnSmemMax) return MR_ERR_UNSUPPORTED;
    MR_CUDA_TRY(cudaMemsetAsync(dh, 0, (size_t)N * H * 4, st), "cudaMemsetAsync(attn dh)");
    MR_CUDA_TRY(cudaMemsetAsync(dprojected, 0, (size_t)N * L * H * 4, st), "cudaMemsetAsync(attn dprojected)");
    MR_CUDA_TRY(cudaMemsetAsync(dmemory, 0, (size_t)N * L * D * 4, st), "cudaMemsetAsync(attn dmemory)");
    MR_CUDA_TRY(cudaMemsetAsync(dv, 0, (size_t)H * 4, st), "cudaMemsetAsync(attn dv)");
    MR_CUDA_TRY(cudaMemsetAsync(dwordtab, 0, (size_t)V * H * 4, st), "cudaMemsetAsync(attn dwordtab)");
    MR_CUDA_TRY(cudaMemsetAsync(sync, 0, 8, st), "cudaMemsetAsync(attn sync)");
    int rc = ensure_dyn_smem((const void *)attn_bwd_kernel, smem, "attn_bwd smem attr");
    if (rc) return rc;
    int max_blocks = 0;
    MR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, attn_bwd_kernel, kAttnThreads, smem), "attn_bwd occupancy");
    if (max_blocks < 1) return MR_ERR_UNSUPPORTED;
    MR_CUDA_TRY(launch_coop(attn_bwd_kernel, dim3((unsigned)grid), kAttnThreads, smem, st, a), "attn_bwd_kernel");
    return check_launch("attn_bwd_kernel");
}

/* error word of a training forward / backward (sync[1]; 0 = fine, 1 = a grid barrier timed out); synchronises the stream */
int mr_attn_sync_status(const void *sync, void *stream, int *status) {
    if (!sync || !status) return MR_ERR_NULL_POINTER;
    unsigned words[2] = {0, 0};
    MR_CUDA_TRY(cudaMemcpyAsync(words, sync, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync(attn sync)");
    MR_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize(attn sync)");
    *status = (int)words[1];
    return MR_OK;
}

}  // extern "C"
