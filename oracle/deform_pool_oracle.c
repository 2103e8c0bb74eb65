/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/ctc2d_oracle.c header for the rules).
 *
 * CPU restatement of MegReader's deformable position-sensitive RoI pooling.
 * The reference has no CPU implementation (assets/ops/dcn/functions/deform_pool.py:32-33,52-53 raise) and its sources
 * do not build against torch 2.11, so this follows the two CUDA kernels statement by statement:
 *
 *   forward   <- assets/ops/dcn/src/deform_pool_cuda_kernel.cu:52-145  (DeformablePSROIPoolForwardKernel)
 *   backward  <- :147-263                                              (DeformablePSROIPoolBackwardAccKernel)
 *   host      <- :265-363 (num_classes / channels_each_class) and deform_pool_cuda.cpp:29-81
 *
 * PARITY UNPINNED: the reference ships no tests, goldens or CPU path for this op and no third-party implementation
 * of the same sampling rule is available offline; tests/test_oracle_deform_pool.py pins the backward to the
 * forward by central finite differences (fp64) and checks closed-form cases (constant maps, integer-aligned bins).
 *
 * Layouts: data [B, C, H, W]; rois [n, 5] = (batch index, x1, y1, x2, y2); trans [n, 2*num_classes, part, part];
 * out / top_count [n, output_dim, P, P].  in_grad / trans_grad are ACCUMULATED (caller zeroes them).
 */
#include <math.h>
#include <stdint.h>

#ifndef REAL
#define REAL double
#define SUF f64
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

typedef struct {
    int batch;                 /* image index of the roi */
    REAL roi_w, roi_h;         /* clamped at 0.1 (:88-89) */
    REAL wstart, hstart;       /* first sample position of the bin, learned shift applied (:105-108) */
    REAL sub_w, sub_h;         /* sample spacing */
    int chan;                  /* position-sensitive input channel (:130) */
    int tx_index, ty_index;    /* flat indices into trans, -1 when no_trans */
} Bin;

/* geometry of output element (n, ctop, ph, pw): identical arithmetic in forward and backward (:71-116, :174-207) */
static Bin bin_geometry(const REAL *rois, const REAL *trans, int n, int ctop, int ph, int pw, int no_trans,
                        REAL spatial_scale, REAL trans_std, int pooled, int part_size, int sample_per_part,
                        int group_size, int num_classes, int channels_each_class) {
    Bin b;
    const REAL *r = rois + (int64_t)n * 5;
    b.batch = (int)r[0];
    const REAL start_w = (REAL)round(r[1]) * spatial_scale - (REAL)0.5;
    const REAL start_h = (REAL)round(r[2]) * spatial_scale - (REAL)0.5;
    const REAL end_w = (REAL)(round(r[3]) + 1.) * spatial_scale - (REAL)0.5;
    const REAL end_h = (REAL)(round(r[4]) + 1.) * spatial_scale - (REAL)0.5;
    b.roi_w = end_w - start_w > (REAL)0.1 ? end_w - start_w : (REAL)0.1;
    b.roi_h = end_h - start_h > (REAL)0.1 ? end_h - start_h : (REAL)0.1;
    const REAL bin_h = b.roi_h / (REAL)pooled, bin_w = b.roi_w / (REAL)pooled;
    b.sub_h = bin_h / (REAL)sample_per_part;
    b.sub_w = bin_w / (REAL)sample_per_part;
    const int part_h = (int)floor((REAL)ph / pooled * part_size);
    const int part_w = (int)floor((REAL)pw / pooled * part_size);
    const int class_id = ctop / channels_each_class;
    REAL tx = 0, ty = 0;
    b.tx_index = b.ty_index = -1;
    if (!no_trans) {
        b.tx_index = (((n * num_classes + class_id) * 2) * part_size + part_h) * part_size + part_w;
        b.ty_index = (((n * num_classes + class_id) * 2 + 1) * part_size + part_h) * part_size + part_w;
        tx = trans[b.tx_index] * trans_std;
        ty = trans[b.ty_index] * trans_std;
    }
    b.wstart = (REAL)pw * bin_w + start_w + tx * b.roi_w;
    b.hstart = (REAL)ph * bin_h + start_h + ty * b.roi_h;
    int gw = (int)floor((REAL)pw * group_size / pooled);
    int gh = (int)floor((REAL)ph * group_size / pooled);
    gw = gw < 0 ? 0 : (gw > group_size - 1 ? group_size - 1 : gw);
    gh = gh < 0 ? 0 : (gh > group_size - 1 ? group_size - 1 : gh);
    b.chan = (ctop * group_size + gh) * group_size + gw;
    return b;
}

void FN(deform_psroi_forward)(const REAL *data, const REAL *rois, const REAL *trans, int channels, int height, int width,
                              int num_rois, int channels_trans, int no_trans, REAL spatial_scale, int output_dim,
                              int group_size, int pooled, int part_size, int sample_per_part, REAL trans_std, REAL *out,
                              REAL *top_count) {
    const int num_classes = no_trans ? 1 : channels_trans / 2;
    const int cec = no_trans ? output_dim : output_dim / num_classes;
    for (int n = 0; n < num_rois; ++n)
        for (int ctop = 0; ctop < output_dim; ++ctop)
            for (int ph = 0; ph < pooled; ++ph)
                for (int pw = 0; pw < pooled; ++pw) {
                    const Bin b = bin_geometry(rois, trans, n, ctop, ph, pw, no_trans, spatial_scale, trans_std, pooled,
                                               part_size, sample_per_part, group_size, num_classes, cec);
                    const REAL *plane = data + ((int64_t)b.batch * channels + b.chan) * height * width;
                    REAL sum = 0;
                    int cnt = 0;
                    for (int ih = 0; ih < sample_per_part; ++ih)
                        for (int iw = 0; iw < sample_per_part; ++iw) {
                            REAL w = b.wstart + iw * b.sub_w, h = b.hstart + ih * b.sub_h;
                            if (w < -0.5 || w > width - 0.5 || h < -0.5 || h > height - 0.5) continue;   /* :127-130 */
                            w = w < 0 ? 0 : (w > width - 1. ? width - 1. : w);
                            h = h < 0 ? 0 : (h > height - 1. ? height - 1. : h);
                            const int x1 = (int)floor(w), x2 = (int)ceil(w), y1 = (int)floor(h), y2 = (int)ceil(h);
                            const REAL dx = w - x1, dy = h - y1;                                        /* :34-50 */
                            sum += (1 - dx) * (1 - dy) * plane[y1 * width + x1] + (1 - dx) * dy * plane[y2 * width + x1] +
                                   dx * (1 - dy) * plane[y1 * width + x2] + dx * dy * plane[y2 * width + x2];
                            ++cnt;
                        }
                    const int64_t o = (((int64_t)n * output_dim + ctop) * pooled + ph) * pooled + pw;
                    out[o] = cnt == 0 ? (REAL)0 : sum / cnt;
                    top_count[o] = (REAL)cnt;
                }
}

void FN(deform_psroi_backward)(const REAL *out_grad, const REAL *data, const REAL *rois, const REAL *trans,
                               const REAL *top_count, int channels, int height, int width, int num_rois,
                               int channels_trans, int no_trans, REAL spatial_scale, int output_dim, int group_size,
                               int pooled, int part_size, int sample_per_part, REAL trans_std, REAL *in_grad,
                               REAL *trans_grad) {
    const int num_classes = no_trans ? 1 : channels_trans / 2;
    const int cec = no_trans ? output_dim : output_dim / num_classes;
    for (int n = 0; n < num_rois; ++n)
        for (int ctop = 0; ctop < output_dim; ++ctop)
            for (int ph = 0; ph < pooled; ++ph)
                for (int pw = 0; pw < pooled; ++pw) {
                    const int64_t o = (((int64_t)n * output_dim + ctop) * pooled + ph) * pooled + pw;
                    if (top_count[o] <= 0) continue;                                                   /* :209-212 */
                    const Bin b = bin_geometry(rois, trans, n, ctop, ph, pw, no_trans, spatial_scale, trans_std, pooled,
                                               part_size, sample_per_part, group_size, num_classes, cec);
                    const REAL diff = out_grad[o] / top_count[o];
                    const int64_t base = ((int64_t)b.batch * channels + b.chan) * height * width;
                    for (int ih = 0; ih < sample_per_part; ++ih)
                        for (int iw = 0; iw < sample_per_part; ++iw) {
                            REAL w = b.wstart + iw * b.sub_w, h = b.hstart + ih * b.sub_h;
                            if (w < -0.5 || w > width - 0.5 || h < -0.5 || h > height - 0.5) continue;
                            w = w < 0 ? 0 : (w > width - 1. ? width - 1. : w);
                            h = h < 0 ? 0 : (h > height - 1. ? height - 1. : h);
                            const int x0 = (int)floor(w), x1 = (int)ceil(w), y0 = (int)floor(h), y1 = (int)ceil(h);
                            const REAL dx = w - x0, dy = h - y0;
                            in_grad[base + y0 * width + x0] += (1 - dx) * (1 - dy) * diff;                /* :236-243 */
                            in_grad[base + y1 * width + x0] += (1 - dx) * dy * diff;
                            in_grad[base + y0 * width + x1] += dx * (1 - dy) * diff;
                            in_grad[base + y1 * width + x1] += dx * dy * diff;
                            if (no_trans) continue;
                            const REAL u00 = data[base + y0 * width + x0], u01 = data[base + y1 * width + x0];
                            const REAL u10 = data[base + y0 * width + x1], u11 = data[base + y1 * width + x1];
                            REAL gx = (u11 * dy + u10 * (1 - dy) - u01 * dy - u00 * (1 - dy)) * trans_std * diff;     /* :253-256 */
                            REAL gy = (u11 * dx + u01 * (1 - dx) - u10 * dx - u00 * (1 - dx)) * trans_std * diff;
                            trans_grad[b.tx_index] += gx * b.roi_w;
                            trans_grad[b.ty_index] += gy * b.roi_h;
                        }
                }
}
