/*
 * megreader_b200 — C-ABI of the B200-native OCR hot path (drop-in for MegReader's native ops).
 *
 * Plain pointers and sizes only: no torch / ATen types.  Every pointer is a DEVICE pointer unless
 * the name ends in `_host`.  `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream, which is what the reference launches on).  Every entry point returns an mr_status
 * (0 = OK); mr_status_string() gives the reference's error text for it.  Nothing here allocates
 * device memory unless stated; outputs are caller-allocated exactly like the reference's ATen
 * tensors (shapes in each comment).  Functions are asynchronous with respect to the host.
 *
 * The reference interface each entry replaces is cited as file:line under /root/reference.
 */
#ifndef MEGREADER_B200_H
#define MEGREADER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MR_OK = 0,
    MR_ERR_NULL_POINTER = 1,
    MR_ERR_BLANK_RANGE = 2,       /* "blank must be in label range"      ctc2d_cuda.cu:40 */
    MR_ERR_TARGET_TOO_LONG = 3,   /* "max target length out of range"    ctc2d_cuda_kernel.cu:220 */
    MR_ERR_BAD_SHAPE = 4,
    MR_ERR_UNSUPPORTED = 5,       /* shape does not fit the on-chip staging of this build */
    MR_ERR_CUDA = 6,              /* a CUDA runtime call failed; see mr_last_cuda_error() */
    MR_ERR_NO_DEVICE = 7
} mr_status;

const char *mr_status_string(int status);
const char *mr_last_cuda_error(void);
/* library/ABI version, bumped when a signature changes */
int mr_abi_version(void);
/* number of kernels this library has launched since load / since the last reset (bench.py's gpu_launches) */
int64_t mr_launch_count(void);
void mr_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * 2D-CTC  (replaces pybind module ops.ctc_2d.ctc_2d_csrc: ops/ctc_2d/csrc/ctc2d.cpp:3-6)
 *
 * log_probs [T,H,N,C] contiguous; targets [N,S] int64 with element strides (tg_stride_n, tg_stride_s);
 * input_lengths, target_lengths [N] int64; blank in [0,C); 2S+1 <= 1024.
 * ---------------------------------------------------------------------------------------------- */

/* ctc2d_forward: ops/ctc_2d/csrc/ctc2d.h:7-21 -> ctc2d_cuda.cu:30-45 -> ctc2d_cuda_kernel.cu:54-251 (K1).
 * Writes nll [N] and log_alpha [N,T,H,2S+1] (every element is written; no pre-zeroing needed).
 * `fast_math` != 0 uses ex2/lg2.approx (f32 only); 0 uses expf/logf. */
int mr_ctc2d_forward_f32(const float *log_probs, const int64_t *targets, const int64_t *input_lengths,
                         const int64_t *target_lengths, int64_t T, int64_t H, int64_t N, int64_t C, int64_t S,
                         int64_t tg_stride_n, int64_t tg_stride_s, int64_t blank, int fast_math,
                         float *nll, float *log_alpha, void *stream);
int mr_ctc2d_forward_f64(const double *log_probs, const int64_t *targets, const int64_t *input_lengths,
                         const int64_t *target_lengths, int64_t T, int64_t H, int64_t N, int64_t C, int64_t S,
                         int64_t tg_stride_n, int64_t tg_stride_s, int64_t blank, int fast_math,
                         double *nll, double *log_alpha, void *stream);

/* ctc2d_backward: ops/ctc_2d/csrc/ctc2d.h:24-43 -> ctc2d_cuda_kernel.cu:520-629 (K2 + K3, is_large = 0).
 * Writes grad [T,H,N,C] (every element).  grad_out [N] with element stride grad_out_stride.
 * `log_alpha` and `nll` are accepted for signature parity; this implementation re-derives both from
 * log_probs on chip (cheaper than re-reading 2S+1 states per pixel from HBM) and ignores them (may be NULL). */
int mr_ctc2d_backward_f32(const float *grad_out, int64_t grad_out_stride, const float *log_probs,
                          const int64_t *targets, const int64_t *input_lengths, const int64_t *target_lengths,
                          const float *nll, const float *log_alpha,
                          int64_t T, int64_t H, int64_t N, int64_t C, int64_t S,
                          int64_t tg_stride_n, int64_t tg_stride_s, int64_t blank, int fast_math,
                          float *grad, void *stream);
int mr_ctc2d_backward_f64(const double *grad_out, int64_t grad_out_stride, const double *log_probs,
                          const int64_t *targets, const int64_t *input_lengths, const int64_t *target_lengths,
                          const double *nll, const double *log_alpha,
                          int64_t T, int64_t H, int64_t N, int64_t C, int64_t S,
                          int64_t tg_stride_n, int64_t tg_stride_s, int64_t blank, int fast_math,
                          double *grad, void *stream);

/* Training pair used by CTCLoss2DFunction (ops/ctc_2d/ctc_loss_2d.py:7-37) when log_probs.requires_grad:
 * the forward keeps no log_alpha; it writes nll [N] and a per-(t,class) factor `gfac` [N,T,C] such that
 *   grad[t,h,b,c] = exp(log_probs[t,h,b,c]) * gfac[b,t,c] * grad_out[b]        (same values as K3),
 * which mr_ctc2d_backward_apply streams out.  Total HBM traffic 3*|log_probs| instead of
 * 3*|log_probs| + 2*|log_alpha|. */
int mr_ctc2d_forward_train_f32(const float *log_probs, const int64_t *targets, const int64_t *input_lengths,
                               const int64_t *target_lengths, int64_t T, int64_t H, int64_t N, int64_t C, int64_t S,
                               int64_t tg_stride_n, int64_t tg_stride_s, int64_t blank, int fast_math,
                               float *nll, float *gfac, void *stream);
int mr_ctc2d_backward_apply_f32(const float *grad_out, int64_t grad_out_stride, const float *log_probs,
                                const float *gfac, int64_t T, int64_t H, int64_t N, int64_t C, int fast_math,
                                float *grad, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MEGREADER_B200_H */
