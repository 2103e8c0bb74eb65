// Greedy decoding loop of the attention recogniser head as ONE persistent kernel (round 2).
//
//   reference: decoders/attention_decoder.py:119-131 (the eval loop: max_size steps of AttentionRNNCell.forward, :187-231, each of
//   them ~15 framework kernels: embedding, Linear, cat + Linear + tanh + bmm for the energies (:160-171), softmax, bmm for the
//   context, GRUCell, Linear, softmax, argmax).
//
// Per step t (h_0 = 0, word_0 = blank):
//   P1  fh[n, :]    = Wa_h . h[n]                                    (the hidden half of the additive-attention Linear; the encoder
//                                                                      half `projected` is step-invariant and computed once by the caller)
//   P2  word[n]     = argmax_v (Wout . h[n] + bout)                   (output of the PREVIOUS step; t = 0: blank)     -> pred[n, t-1]
//       score[l]    = v . tanh(projected[n, l] + fh[n]);  a = softmax_l(score);  context = sum_l a[l] memory[n, l]
//       x[n]        = [wordtab[word[n]] ; context]
//   P3  h'[n, j]    = GRU(x[n], h[n])_j                               (gi = W_ih x + b_ih, gh = W_hh h + b_hh, r, z, n gates)
// and one more P2 head after the last step.  P1 / P3 are parallel over output columns / hidden units: every CTA owns ceil(H / grid)
// of them and keeps ITS rows of Wa_h, W_ih, W_hh in shared memory for the whole loop (weights are read from HBM once per call, not
// once per step); P2 is parallel over samples.  The three phases are separated by grid-wide barriers (monotonic arrival counter,
// cooperative launch so that co-residency is guaranteed; every wait is bounded and raises an error word instead of hanging).
// All arithmetic is fp32 with accurate tanhf / expf: the results match the framework composition to rounding, and the decoded
// strings are identical on the committed goldens.
#include "common.cuh"
#include <math.h>

namespace {
using namespace mr;

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
    float *h0, *h1;           // [N][H] ping-pong hidden state (h0 zero-filled by the caller)
    float *fh;                // [N][H]
    float *x;                 // [N][H + D]
    int *word;                // [N]
    int *pred;                // [N][S]
    float *prob;              // [N][S][V] softmax of the logits (the reference's per-step output), or nullptr
    unsigned *sync;           // [2]: arrival counter (zeroed), error word
    int N, L, H, D, V, S, blank, upc;   // upc = hidden units (and Wa_h columns) per CTA
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
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

constexpr int kAttnThreads = 512;

__global__ void __launch_bounds__(kAttnThreads, 1) attn_decode_kernel(AttnArgs a) {
    extern __shared__ float sm[];
    const int H = a.H, D = a.D, X = a.H + a.D, V = a.V, L = a.L, N = a.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kAttnThreads / 32;
    // hidden units / Wa_h columns of this CTA
    const int j0 = min((int)blockIdx.x * a.upc, H), j1 = min(j0 + a.upc, H), nj = j1 - j0;
    // shared memory: [upc][H] Wa_h rows | [3 upc][X] W_ih rows | [3 upc][H] W_hh rows | scratch of phase 2
    float *s_wa = sm;
    float *s_wih = s_wa + (size_t)a.upc * H;
    float *s_whh = s_wih + (size_t)3 * a.upc * X;
    float *s_p2 = s_whh + (size_t)3 * a.upc * H;       // fh / h row (H), scores (L), logits (V)
    for (int i = threadIdx.x; i < nj * H; i += kAttnThreads) s_wa[i] = a.wa_h[(int64_t)(j0 + i / H) * a.ld_wa + i % H];
    for (int g = 0; g < 3; ++g) {
        for (int i = threadIdx.x; i < nj * X; i += kAttnThreads) s_wih[(size_t)(g * a.upc) * X + i] = a.w_ih[((int64_t)g * H + j0) * X + i];
        for (int i = threadIdx.x; i < nj * H; i += kAttnThreads) s_whh[(size_t)(g * a.upc) * H + i] = a.w_hh[((int64_t)g * H + j0) * H + i];
    }
    __syncthreads();
    float *s_row = s_p2, *s_score = s_p2 + H, *s_logit = s_score + L;
    __shared__ int s_word;
    unsigned bar = 0;
    const float *h = a.h0;
    float *hn = a.h1;

    for (int t = 0; t <= a.S; ++t) {
        // ---------------- P1: fh = Wa_h . h (columns j0..j1 of every sample); skipped after the last step
        if (t < a.S) {
            for (int n = warp; n < N; n += nwarps) {
                const float *hr = h + (int64_t)n * H;
                for (int cb = 0; cb < nj; cb += 4) {                  // four columns at a time share the loads of h
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int k = lane; k < H; k += 32) {
                        const float hv = hr[k];
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[c] = fmaf(s_wa[min(cb + c, nj - 1) * H + k], hv, acc[c]);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float r = warp_sum(acc[c]);
                        if (lane == 0 && cb + c < nj) a.fh[(int64_t)n * H + j0 + cb + c] = r;
                    }
                }
            }
            grid_barrier(a.sync, bar++);
        }
        // ---------------- P2: per sample -- previous step's output symbol, attention, context, GRU input
        for (int n = blockIdx.x; n < N; n += gridDim.x) {
            __syncthreads();
            for (int i = threadIdx.x; i < H; i += kAttnThreads) s_row[i] = h[(int64_t)n * H + i];
            __syncthreads();
            if (t > 0) {
                for (int vv = warp; vv < V; vv += nwarps) {
                    float acc = 0.f;
                    for (int k = lane; k < H; k += 32) acc = fmaf(a.w_out[(int64_t)vv * H + k], s_row[k], acc);
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
                        const float ob = __shfl_xor_sync(0xffffffffu, best, s);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }       // first maximum, like torch.argmax
                    }
                    if (lane == 0) { s_word = bi; a.pred[(int64_t)n * a.S + t - 1] = bi; a.word[n] = bi; }
                    if (a.prob) {
                        float den = 0.f;
                        for (int vv = lane; vv < V; vv += 32) den += expf(s_logit[vv] - best);
                        den = warp_sum(den);
                        for (int vv = lane; vv < V; vv += 32) a.prob[((int64_t)n * a.S + t - 1) * V + vv] = expf(s_logit[vv] - best) / den;
                    }
                }
            } else if (threadIdx.x == 0) {
                s_word = a.blank;
            }
            if (t == a.S) continue;
            __syncthreads();
            for (int i = threadIdx.x; i < H; i += kAttnThreads) s_row[i] = a.fh[(int64_t)n * H + i];
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
                for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
                float den = 0.f;
                for (int l = lane; l < L; l += 32) den += expf(s_score[l] - mx);
                den = warp_sum(den);
                for (int l = lane; l < L; l += 32) s_score[l] = expf(s_score[l] - mx) / den;
            }
            __syncthreads();
            const float *mem = a.memory + (int64_t)n * L * D;
            float *xr = a.x + (int64_t)n * X;
            for (int d = threadIdx.x; d < D; d += kAttnThreads) {
                float acc = 0.f;
                for (int l = 0; l < L; ++l) acc = fmaf(s_score[l], mem[(int64_t)l * D + d], acc);
                xr[H + d] = acc;
            }
            const float *wt = a.wordtab + (int64_t)s_word * H;
            for (int i = threadIdx.x; i < H; i += kAttnThreads) xr[i] = wt[i];
        }
        if (t == a.S) break;
        grid_barrier(a.sync, bar++);
        // ---------------- P3: GRU cell for hidden units j0..j1 of every sample (torch.nn.GRUCell gate order r, z, n)
        for (int n = warp; n < N; n += nwarps) {
            const float *xr = a.x + (int64_t)n * X, *hr = h + (int64_t)n * H;
            for (int cb = 0; cb < nj; cb += 4) {                      // four hidden units (12 gate rows) share the loads of x and h
                float ai[3][4], ah[3][4];
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int c = 0; c < 4; ++c) ai[g][c] = ah[g][c] = 0.f;
                for (int k = lane; k < X; k += 32) {
                    const float xv = xr[k];
#pragma unroll
                    for (int g = 0; g < 3; ++g)
#pragma unroll
                        for (int c = 0; c < 4; ++c) ai[g][c] = fmaf(s_wih[(size_t)(g * a.upc + min(cb + c, nj - 1)) * X + k], xv, ai[g][c]);
                }
                for (int k = lane; k < H; k += 32) {
                    const float hv = hr[k];
#pragma unroll
                    for (int g = 0; g < 3; ++g)
#pragma unroll
                        for (int c = 0; c < 4; ++c) ah[g][c] = fmaf(s_whh[(size_t)(g * a.upc + min(cb + c, nj - 1)) * H + k], hv, ah[g][c]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float gi[3], gh[3];
#pragma unroll
                    for (int g = 0; g < 3; ++g) { gi[g] = warp_sum(ai[g][c]); gh[g] = warp_sum(ah[g][c]); }
                    if (lane == 0 && cb + c < nj) {
                        const int j = j0 + cb + c;
                        const float r = 1.f / (1.f + expf(-(gi[0] + a.b_ih[j] + gh[0] + a.b_hh[j])));
                        const float z = 1.f / (1.f + expf(-(gi[1] + a.b_ih[H + j] + gh[1] + a.b_hh[H + j])));
                        const float nn = tanhf(gi[2] + a.b_ih[2 * H + j] + r * (gh[2] + a.b_hh[2 * H + j]));
                        hn[(int64_t)n * H + j] = (1.f - z) * nn + z * hr[j];
                    }
                }
            }
        }
        grid_barrier(a.sync, bar++);
        const float *tmp = h; h = hn; hn = const_cast<float *>(tmp);
    }
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

}  // namespace

extern "C" {

/* floats of scratch the decode loop needs: h (2 x N x H), fh (N x H), x (N x (2H + E)), word (N ints), sync (2 words, 256 B) */
int64_t mr_attn_decode_workspace_bytes(int64_t N, int64_t H, int64_t E) {
    return round_up(N * H * 4, 256) * 3 + round_up(N * (2 * H + E) * 4, 256) + round_up(N * 4, 256) + 256;
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
    AttnArgs a;
    a.projected = projected; a.memory = memory; a.wa_h = wa_h; a.ld_wa = ld_wa; a.v = v; a.wordtab = wordtab;
    a.w_ih = w_ih; a.b_ih = b_ih; a.w_hh = w_hh; a.b_hh = b_hh; a.w_out = w_out; a.b_out = b_out;
    a.N = N; a.L = L; a.H = H; a.D = H + E; a.V = V; a.S = S; a.blank = blank; a.pred = pred; a.prob = prob;
    unsigned char *ws = (unsigned char *)workspace;
    const int64_t hb = round_up((int64_t)N * H * 4, 256), xb = round_up((int64_t)N * (2 * H + E) * 4, 256), wb = round_up((int64_t)N * 4, 256);
    a.h0 = (float *)ws; a.h1 = (float *)(ws + hb); a.fh = (float *)(ws + 2 * hb); a.x = (float *)(ws + 3 * hb);
    a.word = (int *)(ws + 3 * hb + xb); a.sync = (unsigned *)(ws + 3 * hb + xb + wb);
    MR_CUDA_TRY(cudaMemsetAsync(a.h0, 0, (size_t)N * H * 4, st), "cudaMemsetAsync(attn h0)");
    MR_CUDA_TRY(cudaMemsetAsync(a.sync, 0, 256, st), "cudaMemsetAsync(attn sync)");
    const int X = 2 * H + E;
    int grid = sm_count();
    int upc = (int)ceil_div(H, grid);
    auto smem_of = [&](int u) { return ((size_t)u * H + (size_t)3 * u * X + (size_t)3 * u * H + H + L + V) * sizeof(float); };
    if (smem_of(upc) > 220 * 1024) return MR_ERR_UNSUPPORTED;
    const size_t smem = smem_of(upc);
    a.upc = upc;
    int rc = ensure_dyn_smem((const void *)attn_decode_kernel, (int)smem, "attn_decode smem attr");
    if (rc) return rc;
    int max_blocks = 0;
    MR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, attn_decode_kernel, kAttnThreads, smem), "attn_decode occupancy");
    if (max_blocks < 1) return MR_ERR_UNSUPPORTED;
    MR_CUDA_TRY(launch_coop(attn_decode_kernel, dim3((unsigned)grid), kAttnThreads, smem, st, a), "attn_decode_kernel");
    return check_launch("attn_decode_kernel");
}

/* error word of the last decode on this workspace (0 = fine, 1 = a grid barrier timed out); synchronises the stream */
int mr_attn_decode_status(const void *workspace, int64_t N, int64_t H, int64_t E, void *stream, int *status) {
    if (!workspace || !status) return MR_ERR_NULL_POINTER;
    const unsigned char *ws = (const unsigned char *)workspace;
    const int64_t hb = round_up(N * H * 4, 256), xb = round_up(N * (2 * H + E) * 4, 256), wb = round_up(N * 4, 256);
    unsigned words[2] = {0, 0};
    MR_CUDA_TRY(cudaMemcpyAsync(words, ws + 3 * hb + xb + wb, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync(attn status)");
    MR_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize(attn status)");
    *status = (int)words[1];
    return MR_OK;
}

}  // extern "C"
