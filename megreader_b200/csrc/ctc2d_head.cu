// Fused epilogue of the 2D-CTC head (decoders/ctc_decoder2d.py:37-45):
//
//     mask     = softmax_H(mask_logits)          [N,1,H,W]      (nn.Softmax(dim=2) at the end of pred_mask, :21)
//     classify = softmax_C(classify_logits)      [N,C,H,W]      (:41)
//     pred     = log(max(mask * classify, tiny)) [N,C,H,W]      (:43-44)
//     log_probs = pred.permute(3,2,0,1).contiguous()   -> [W,H,N,C] = the (T,H,N,C) operand of ctc_loss_2d   (:45)
//
// The reference runs this as two softmax kernels, a product, a max, a log and a permuting copy: 6 passes over
// N*C*H*W floats.  Here: ONE pass forward (read the NCHW logits once, write (T,H,N,C) once) and one pass backward
// (+ a tiny per-column fix-up for the mask branch).  The backward can take the 2D-CTC training factor `gfac`
// directly (grad[t,h,n,c] = exp(log_probs) * gfac[t,n,c] * grad_out[n], include/megreader_b200.h) so that the loss
// gradient w.r.t. log_probs never exists in HBM.
//
// Tiling: a block owns (NT samples) x (one height) x (32 columns); the NCHW side is read / written as 128-byte rows
// along W, the (T,H,N,C) side as NT*C contiguous floats per column, with a padded shared-memory tile in between.
// HBM-bound: algorithmic bytes per sample = 2*C*H*W*4 (+ mask) forward, 3*C*H*W*4 backward with explicit grad,
// 2*C*H*W*4 + T*C*4 with gfac.
#include "common.cuh"
#include <math.h>

namespace {
using namespace mr;

constexpr int kTW = 32;          // columns per block
constexpr int kThreads = 256;

struct HeadGeo { int N, C, H, W, NT; };

// log-softmax of the mask logits over H at (n, w) for height h; also returns softmax value
__device__ __forceinline__ float mask_logsoftmax(const float *__restrict__ m, int n, int h, int w, int H, int W) {
    const float *col = m + (int64_t)n * H * W + w;
    float mx = -INFINITY;
    for (int k = 0; k < H; ++k) mx = fmaxf(mx, __ldg(col + (int64_t)k * W));
    float s = 0.f;
    for (int k = 0; k < H; ++k) s += expf(__ldg(col + (int64_t)k * W) - mx);
    return __ldg(col + (int64_t)h * W) - mx - logf(s);
}

// tile[nl][c][w] <-> z[n0+nl, c, h, w0+w]  (rows of 32 floats along W)
__device__ __forceinline__ void load_nchw_tile(const HeadGeo &g, const float *__restrict__ z, int n0, int h, int w0,
                                               float *tile) {
    const int w = threadIdx.x & 31, r0 = threadIdx.x >> 5;
    const int rows = g.NT * g.C;
    for (int r = r0; r < rows; r += kThreads / 32) {
        const int nl = r / g.C, c = r - nl * g.C;
        const int n = n0 + nl;
        float v = 0.f;
        if (n < g.N && w0 + w < g.W) v = __ldg(z + (((int64_t)n * g.C + c) * g.H + h) * g.W + w0 + w);
        tile[r * (kTW + 1) + w] = v;
    }
}

__device__ __forceinline__ void store_nchw_tile(const HeadGeo &g, float *__restrict__ z, int n0, int h, int w0,
                                                const float *tile) {
    const int w = threadIdx.x & 31, r0 = threadIdx.x >> 5;
    const int rows = g.NT * g.C;
    for (int r = r0; r < rows; r += kThreads / 32) {
        const int nl = r / g.C, c = r - nl * g.C;
        const int n = n0 + nl;
        if (n < g.N && w0 + w < g.W) z[(((int64_t)n * g.C + c) * g.H + h) * g.W + w0 + w] = tile[r * (kTW + 1) + w];
    }
}

// forward: logits -> log_probs (T=W, H, N, C)
__global__ void __launch_bounds__(kThreads)
ctc2d_head_fwd_kernel(HeadGeo g, const float *__restrict__ mask_logits, const float *__restrict__ cls_logits,
                      float log_tiny, float *__restrict__ lp) {
    extern __shared__ float tile[];                      // [NT*C][33]
    const int w0 = blockIdx.x * kTW, h = blockIdx.y, n0 = blockIdx.z * g.NT;
    load_nchw_tile(g, cls_logits, n0, h, w0, tile);
    __syncthreads();
    // one thread per (sample, column): class log-softmax + mask log-softmax, clamp at log(tiny)
    for (int item = threadIdx.x; item < g.NT * kTW; item += kThreads) {
        const int nl = item >> 5, w = item & 31;
        const int n = n0 + nl;
        if (n >= g.N || w0 + w >= g.W) continue;
        float *col = tile + (nl * g.C) * (kTW + 1) + w;
        float mx = -INFINITY;
        for (int c = 0; c < g.C; ++c) mx = fmaxf(mx, col[c * (kTW + 1)]);
        float s = 0.f;
        for (int c = 0; c < g.C; ++c) s += expf(col[c * (kTW + 1)] - mx);
        const float shift = mask_logsoftmax(mask_logits, n, h, w0 + w, g.H, g.W) - mx - logf(s);
        for (int c = 0; c < g.C; ++c) col[c * (kTW + 1)] = fmaxf(col[c * (kTW + 1)] + shift, log_tiny);
    }
    __syncthreads();
    // (T,H,N,C): for each column the NT*C values of this block are contiguous
    const int per_col = g.NT * g.C;
    const int nvalid = min(g.NT, g.N - n0) * g.C;
    for (int w = 0; w < kTW && w0 + w < g.W; ++w) {
        float *dst = lp + (((int64_t)(w0 + w) * g.H + h) * g.N + n0) * g.C;
        for (int e = threadIdx.x; e < per_col; e += kThreads)
            if (e < nvalid) dst[e] = tile[e * (kTW + 1) + w];
    }
}

// backward: d(log_probs) [or gfac/grad_out] -> d(classify logits) [N,C,H,W], and G[n,h,w] = sum_c g  (mask branch)
template <bool FACTORED>
__global__ void __launch_bounds__(kThreads)
ctc2d_head_bwd_kernel(HeadGeo g, const float *__restrict__ mask_logits, const float *__restrict__ cls_logits,
                      const float *__restrict__ dlp, const float *__restrict__ gfac, const float *__restrict__ go,
                      int64_t go_stride, float log_tiny, float *__restrict__ dcls, float *__restrict__ gsum) {
    extern __shared__ float smem[];
    const int per_col = g.NT * g.C;
    float *zt = smem;                                    // [NT*C][33] logits -> d logits
    float *gt = smem + per_col * (kTW + 1);              // [NT*C][33] upstream gradient (or factor)
    const int w0 = blockIdx.x * kTW, h = blockIdx.y, n0 = blockIdx.z * g.NT;
    load_nchw_tile(g, cls_logits, n0, h, w0, zt);
    const int nvalid = min(g.NT, g.N - n0) * g.C;
    for (int w = 0; w < kTW && w0 + w < g.W; ++w) {
        // explicit gradient is [T,H,N,C]; the factor is [T,N,C] (shared by all heights)
        const float *src = FACTORED ? gfac + ((int64_t)(w0 + w) * g.N + n0) * g.C
                                    : dlp + (((int64_t)(w0 + w) * g.H + h) * g.N + n0) * g.C;
        for (int e = threadIdx.x; e < per_col; e += kThreads) gt[e * (kTW + 1) + w] = e < nvalid ? __ldg(src + e) : 0.f;
    }
    __syncthreads();
    for (int item = threadIdx.x; item < g.NT * kTW; item += kThreads) {
        const int nl = item >> 5, w = item & 31;
        const int n = n0 + nl;
        if (n >= g.N || w0 + w >= g.W) continue;
        float *zc = zt + (nl * g.C) * (kTW + 1) + w;
        const float *gc = gt + (nl * g.C) * (kTW + 1) + w;
        float mx = -INFINITY;
        for (int c = 0; c < g.C; ++c) mx = fmaxf(mx, zc[c * (kTW + 1)]);
        float s = 0.f;
        for (int c = 0; c < g.C; ++c) s += expf(zc[c * (kTW + 1)] - mx);
        const float lse = mx + logf(s);
        const float lm = mask_logsoftmax(mask_logits, n, h, w0 + w, g.H, g.W);
        const float gout = FACTORED ? __ldg(go + (int64_t)n * go_stride) : 1.f;
        float tot = 0.f;
        for (int c = 0; c < g.C; ++c) {
            const float lc = zc[c * (kTW + 1)] - lse;
            const float l = lm + lc;
            float gv = 0.f;
            if (l > log_tiny) gv = FACTORED ? expf(l) * gc[c * (kTW + 1)] * gout : gc[c * (kTW + 1)];   // max(.,tiny): no gradient when clamped
            tot += gv;
            zc[c * (kTW + 1)] = lc;                       // keep log-softmax, combine below
            const_cast<float *>(gc)[c * (kTW + 1)] = gv;
        }
        for (int c = 0; c < g.C; ++c) zc[c * (kTW + 1)] = gc[c * (kTW + 1)] - expf(zc[c * (kTW + 1)]) * tot;
        gsum[((int64_t)n * g.H + h) * g.W + w0 + w] = tot;
    }
    __syncthreads();
    store_nchw_tile(g, dcls, n0, h, w0, zt);
}

// mask branch: dm[n,h,w] = G[n,h,w] - softmax_H(m)[n,h,w] * sum_h' G[n,h',w]   (in place on G)
__global__ void ctc2d_head_mask_bwd_kernel(int N, int H, int W, const float *__restrict__ mask_logits, float *__restrict__ gm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * W) return;
    const int n = (int)(i / W), w = (int)(i - (int64_t)n * W);
    const float *m = mask_logits + (int64_t)n * H * W + w;
    float *gp = gm + (int64_t)n * H * W + w;
    float mx = -INFINITY, tot = 0.f;
    for (int k = 0; k < H; ++k) { mx = fmaxf(mx, m[(int64_t)k * W]); tot += gp[(int64_t)k * W]; }
    float s = 0.f;
    for (int k = 0; k < H; ++k) s += expf(m[(int64_t)k * W] - mx);
    const float inv = 1.f / s;
    for (int k = 0; k < H; ++k) gp[(int64_t)k * W] -= expf(m[(int64_t)k * W] - mx) * inv * tot;
}

int pick_nt(int C, int tiles, size_t *smem) {
    // samples per block: as many as fit 48 KB per tile (two tiles in the backward), at most 8
    int nt = 8;
    while (nt > 1 && (size_t)nt * C * (kTW + 1) * sizeof(float) * tiles > 96 * 1024) nt >>= 1;
    *smem = (size_t)nt * C * (kTW + 1) * sizeof(float) * tiles;
    return *smem <= 200 * 1024 ? nt : 0;
}

}  // namespace

extern "C" {

int mr_ctc2d_head_fwd_f32(const float *mask_logits, const float *cls_logits, int N, int C, int H, int W, float tiny,
                          float *log_probs, void *stream) {
    if (N < 0 || C <= 0 || H <= 0 || W <= 0 || !(tiny > 0.f)) return MR_ERR_BAD_SHAPE;
    if (N == 0) return MR_OK;
    if (!mask_logits || !cls_logits || !log_probs) return MR_ERR_NULL_POINTER;
    HeadGeo g{N, C, H, W, 0};
    size_t smem;
    g.NT = pick_nt(C, 1, &smem);
    if (!g.NT || H > 65535 || ceil_div(N, g.NT) > 65535) return MR_ERR_UNSUPPORTED;
    static size_t attr = 0;
    if (smem > attr) {
        MR_CUDA_TRY(cudaFuncSetAttribute(ctc2d_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "head fwd smem");
        attr = smem;
    }
    dim3 grid((unsigned)ceil_div(W, kTW), (unsigned)H, (unsigned)ceil_div(N, g.NT));
    ctc2d_head_fwd_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(g, mask_logits, cls_logits, logf(tiny), log_probs);
    return check_launch("ctc2d_head_fwd_kernel");
}

/* Exactly one of grad_log_probs [W,H,N,C] or (gfac [W,N,C], grad_out [N] with element stride) is given. */
int mr_ctc2d_head_bwd_f32(const float *mask_logits, const float *cls_logits, const float *grad_log_probs, const float *gfac,
                          const float *grad_out, int64_t grad_out_stride, int N, int C, int H, int W, float tiny,
                          float *grad_cls_logits, float *grad_mask_logits, void *stream) {
    if (N < 0 || C <= 0 || H <= 0 || W <= 0 || !(tiny > 0.f)) return MR_ERR_BAD_SHAPE;
    if (N == 0) return MR_OK;
    if (!mask_logits || !cls_logits || !grad_cls_logits || !grad_mask_logits) return MR_ERR_NULL_POINTER;
    const bool factored = grad_log_probs == nullptr;
    if (factored && (!gfac || !grad_out)) return MR_ERR_NULL_POINTER;
    HeadGeo g{N, C, H, W, 0};
    size_t smem;
    g.NT = pick_nt(C, 2, &smem);
    if (!g.NT || H > 65535 || ceil_div(N, g.NT) > 65535) return MR_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    static size_t attr[2] = {0, 0};
    if (smem > attr[factored]) {
        if (factored) MR_CUDA_TRY(cudaFuncSetAttribute(ctc2d_head_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "head bwd smem");
        else MR_CUDA_TRY(cudaFuncSetAttribute(ctc2d_head_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "head bwd smem");
        attr[factored] = smem;
    }
    dim3 grid((unsigned)ceil_div(W, kTW), (unsigned)H, (unsigned)ceil_div(N, g.NT));
    if (factored)
        ctc2d_head_bwd_kernel<true><<<grid, kThreads, smem, st>>>(g, mask_logits, cls_logits, nullptr, gfac, grad_out, grad_out_stride,
                                                                   logf(tiny), grad_cls_logits, grad_mask_logits);
    else
        ctc2d_head_bwd_kernel<false><<<grid, kThreads, smem, st>>>(g, mask_logits, cls_logits, grad_log_probs, nullptr, nullptr, 0,
                                                                    logf(tiny), grad_cls_logits, grad_mask_logits);
    int rc = check_launch("ctc2d_head_bwd_kernel");
    if (rc) return rc;
    const int64_t cols = (int64_t)N * W;
    ctc2d_head_mask_bwd_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, st>>>(N, H, W, mask_logits, grad_mask_logits);
    return check_launch("ctc2d_head_mask_bwd_kernel");
}

}  // extern "C"
