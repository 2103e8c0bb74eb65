/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/ctc2d_oracle.c header for the rules).
 *
 * CPU restatement of MegReader's deformable convolution (DCNv1 / DCNv2) CUDA op.  The reference has no CPU
 * implementation (assets/ops/dcn/functions/deform_conv.py:40-41,130-131 raise) and its sources do not build
 * against torch 2.11, so this follows the CUDA kernels and the host orchestration statement by statement:
 *
 *   bilinear            <- assets/ops/dcn/src/deform_conv_cuda_kernel.cu:466-496  (dmcn_im2col_bilinear)
 *   grad weight         <- :498-525   (dmcn_get_gradient_weight)
 *   coord weight        <- :527-567   (dmcn_get_coordinate_weight)
 *   im2col              <- :569-632   (K8; K5 :189-242 is the same without mask)
 *   col2im              <- :634-692   (K9; K6 :278-334 without mask)
 *   col2im_coord        <- :694-766   (K10; K7 :372-435 without mask)
 *   forward / backward  <- assets/ops/dcn/src/deform_conv_cuda.cpp:486-564 / :566-679 (per-sample loop, GEMMs)
 *
 * Parity pin: no reference tests exist.  Pinned against torchvision.ops.deform_conv2d (same mmdetection
 * lineage, third-party) where offset spatial size == output size (tests/test_oracle_dcn.py); the flat
 * (Ho,Wo) re-indexing of larger offset maps is restated literally.
 *
 * offset [B, 2*kh*kw*dg, ...] and mask [B, kh*kw*dg, ...] are addressed per sample as base + b*bstride and then
 * FLAT with (Ho, Wo) strides, exactly like the kernels (:599-609).  mask == NULL -> DCNv1 (mask = 1).
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

typedef struct {
    int B, C, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, group, dg, Ho, Wo;
} DcnGeo;

static REAL bilinear(const REAL *im, int data_width, int height, int width, REAL h, REAL w) {
    int h_low = (int)floor((double)h), w_low = (int)floor((double)w);
    int h_high = h_low + 1, w_high = w_low + 1;
    REAL lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
    REAL v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (h_low >= 0 && w_low >= 0) v1 = im[h_low * data_width + w_low];
    if (h_low >= 0 && w_high <= width - 1) v2 = im[h_low * data_width + w_high];
    if (h_high <= height - 1 && w_low >= 0) v3 = im[h_high * data_width + w_low];
    if (h_high <= height - 1 && w_high <= width - 1) v4 = im[h_high * data_width + w_high];
    REAL w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

static REAL gradient_weight(REAL ah, REAL aw, int h, int w, int height, int width) {
    if (ah <= -1 || ah >= height || aw <= -1 || aw >= width) return 0;
    int hl = (int)floor((double)ah), wl = (int)floor((double)aw), hh = hl + 1, wh = wl + 1;
    REAL weight = 0;
    if (h == hl && w == wl) weight = (h + 1 - ah) * (w + 1 - aw);
    if (h == hl && w == wh) weight = (h + 1 - ah) * (aw + 1 - w);
    if (h == hh && w == wl) weight = (ah + 1 - h) * (w + 1 - aw);
    if (h == hh && w == wh) weight = (ah + 1 - h) * (aw + 1 - w);
    return weight;
}

static REAL coordinate_weight(REAL ah, REAL aw, int height, int width, const REAL *im, int data_width, int bp_dir) {
    if (ah <= -1 || ah >= height || aw <= -1 || aw >= width) return 0;
    int hl = (int)floor((double)ah), wl = (int)floor((double)aw), hh = hl + 1, wh = wl + 1;
    REAL weight = 0;
    if (bp_dir == 0) {
        if (hl >= 0 && wl >= 0) weight += -1 * (wl + 1 - aw) * im[hl * data_width + wl];
        if (hl >= 0 && wh <= width - 1) weight += -1 * (aw - wl) * im[hl * data_width + wh];
        if (hh <= height - 1 && wl >= 0) weight += (wl + 1 - aw) * im[hh * data_width + wl];
        if (hh <= height - 1 && wh <= width - 1) weight += (aw - wl) * im[hh * data_width + wh];
    } else {
        if (hl >= 0 && wl >= 0) weight += -1 * (hl + 1 - ah) * im[hl * data_width + wl];
        if (hl >= 0 && wh <= width - 1) weight += (hl + 1 - ah) * im[hl * data_width + wh];
        if (hh <= height - 1 && wl >= 0) weight += -1 * (ah - hl) * im[hh * data_width + wl];
        if (hh <= height - 1 && wh <= width - 1) weight += (ah - hl) * im[hh * data_width + wh];
    }
    return weight;
}

/* K8 for one sample (batch_size = 1 per launch, deform_conv_cuda.cpp:534-538).  col [C*kh*kw, Ho*Wo]. */
static void im2col_one(const DcnGeo *g, const REAL *im, const REAL *off, const REAL *msk, REAL *col) {
    const int K = g->kh * g->kw, cpg = g->C / g->dg, P = g->Ho * g->Wo;
    for (int c = 0; c < g->C; ++c) {
        const int dgi = c / cpg;
        const REAL *imc = im + (size_t)c * g->H * g->W;
        const REAL *offp = off + (size_t)dgi * 2 * K * P;
        const REAL *mskp = msk ? msk + (size_t)dgi * K * P : NULL;
        for (int ho = 0; ho < g->Ho; ++ho)
            for (int wo = 0; wo < g->Wo; ++wo) {
                const int h_in = ho * g->sh - g->ph, w_in = wo * g->sw - g->pw;
                for (int i = 0; i < g->kh; ++i)
                    for (int j = 0; j < g->kw; ++j) {
                        const int k = i * g->kw + j;
                        const REAL oh = offp[((2 * k) * g->Ho + ho) * g->Wo + wo];
                        const REAL ow = offp[((2 * k + 1) * g->Ho + ho) * g->Wo + wo];
                        const REAL m = mskp ? mskp[(k * g->Ho + ho) * g->Wo + wo] : (REAL)1;
                        const REAL h_im = h_in + i * g->dh + oh, w_im = w_in + j * g->dw + ow;
                        REAL val = 0;
                        if (h_im > -1 && w_im > -1 && h_im < g->H && w_im < g->W)
                            val = bilinear(imc, g->W, g->H, g->W, h_im, w_im);
                        col[((size_t)c * K + k) * P + ho * g->Wo + wo] = val * m;
                    }
            }
    }
}

/* K9 for one sample: grad_im [C,H,W] += scatter(col * mask) */
static void col2im_one(const DcnGeo *g, const REAL *col, const REAL *off, const REAL *msk, REAL *grad_im) {
    const int K = g->kh * g->kw, cpg = g->C / g->dg, P = g->Ho * g->Wo;
    for (int c = 0; c < g->C; ++c) {
        const int dgi = c / cpg;
        const REAL *offp = off + (size_t)dgi * 2 * K * P;
        const REAL *mskp = msk ? msk + (size_t)dgi * K * P : NULL;
        for (int i = 0; i < g->kh; ++i)
            for (int j = 0; j < g->kw; ++j)
                for (int ho = 0; ho < g->Ho; ++ho)
                    for (int wo = 0; wo < g->Wo; ++wo) {
                        const int k = i * g->kw + j;
                        const int w_in = wo * g->sw - g->pw, h_in = ho * g->sh - g->ph;
                        const REAL oh = offp[((2 * k) * g->Ho + ho) * g->Wo + wo];
                        const REAL ow = offp[((2 * k + 1) * g->Ho + ho) * g->Wo + wo];
                        const REAL m = mskp ? mskp[(k * g->Ho + ho) * g->Wo + wo] : (REAL)1;
                        const REAL ch = h_in + i * g->dh + oh, cw = w_in + j * g->dw + ow;
                        const REAL top = col[((size_t)c * K + k) * P + ho * g->Wo + wo] * m;
                        const int cur_h = (int)ch, cur_w = (int)cw; /* truncation toward zero (:674-675) */
                        for (int dy = -2; dy <= 2; ++dy)
                            for (int dx = -2; dx <= 2; ++dx)
                                if (cur_h + dy >= 0 && cur_h + dy < g->H && cur_w + dx >= 0 && cur_w + dx < g->W &&
                                    fabs((double)(ch - (cur_h + dy))) < 1 && fabs((double)(cw - (cur_w + dx))) < 1) {
                                    const REAL wgt = gradient_weight(ch, cw, cur_h + dy, cur_w + dx, g->H, g->W);
                                    grad_im[((size_t)c * g->H + cur_h + dy) * g->W + cur_w + dx] += wgt * top;
                                }
                    }
    }
}

/* K10 for one sample: grad_off [2*K*dg, Ho*Wo] (flat), grad_msk [K*dg, Ho*Wo] (flat, may be NULL) */
static void col2im_coord_one(const DcnGeo *g, const REAL *col, const REAL *im, const REAL *off, const REAL *msk,
                             REAL *grad_off, REAL *grad_msk) {
    const int K = g->kh * g->kw, P = g->Ho * g->Wo;
    const int cpg_col = g->C * K / g->dg; /* channel_per_deformable_group as passed to K10 (:850) */
    const int offset_channels = 2 * K * g->dg;
    for (int c = 0; c < offset_channels; ++c)
        for (int h = 0; h < g->Ho; ++h)
            for (int w = 0; w < g->Wo; ++w) {
                REAL val = 0, mval = 0;
                const int dgi = c / (2 * K);
                const REAL *colp = col + (size_t)dgi * cpg_col * P;
                const REAL *imp = im + (size_t)dgi * (cpg_col / K) * g->H * g->W;
                const REAL *offp = off + (size_t)dgi * 2 * K * P;
                const REAL *mskp = msk ? msk + (size_t)dgi * K * P : NULL;
                const int offset_c = c - dgi * 2 * K;
                int cnt = 0;
                for (int col_c = offset_c / 2; col_c < cpg_col; col_c += K) {
                    const int col_pos = (col_c * g->Ho + h) * g->Wo + w;
                    const int bp_dir = offset_c % 2;
                    const int j = (col_pos / g->Wo / g->Ho) % g->kw;
                    const int i = (col_pos / g->Wo / g->Ho / g->kw) % g->kh;
                    const int w_in = w * g->sw - g->pw, h_in = h * g->sh - g->ph;
                    const int k = i * g->kw + j;
                    const REAL oh = offp[((2 * k) * g->Ho + h) * g->Wo + w];
                    const REAL ow = offp[((2 * k + 1) * g->Ho + h) * g->Wo + w];
                    const REAL m = mskp ? mskp[(k * g->Ho + h) * g->Wo + w] : (REAL)1;
                    REAL inv_h = h_in + i * g->dh + oh, inv_w = w_in + j * g->dw + ow;
                    if (inv_h <= -1 || inv_w <= -1 || inv_h >= g->H || inv_w >= g->W) inv_h = inv_w = -2;
                    else mval += colp[col_pos] * bilinear(imp + (size_t)cnt * g->H * g->W, g->W, g->H, g->W, inv_h, inv_w);
                    const REAL wgt = coordinate_weight(inv_h, inv_w, g->H, g->W, imp + (size_t)cnt * g->H * g->W, g->W, bp_dir);
                    val += wgt * colp[col_pos] * m;
                    cnt += 1;
                }
                grad_off[((size_t)c * g->Ho + h) * g->Wo + w] = val;
                if (grad_msk && offset_c % 2 == 0)
                    grad_msk[(((size_t)dgi * K + offset_c / 2) * g->Ho + h) * g->Wo + w] = mval;
            }
}

static void geo_fill(DcnGeo *g, int B, int C, int H, int W, int Cout, int kh, int kw, int sh, int sw, int ph, int pw,
                     int dh, int dw, int group, int dg) {
    g->B = B; g->C = C; g->H = H; g->W = W; g->Cout = Cout; g->kh = kh; g->kw = kw; g->sh = sh; g->sw = sw;
    g->ph = ph; g->pw = pw; g->dh = dh; g->dw = dw; g->group = group; g->dg = dg;
    g->Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
    g->Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
}

/* exposed for unit tests of the gather alone */
void FN(dcn_im2col)(const REAL *input, const REAL *offset, int64_t off_bstride, const REAL *mask, int64_t mask_bstride,
                    int B, int C, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int dg,
                    REAL *columns /* [B][C*kh*kw][Ho*Wo] */) {
    DcnGeo g; geo_fill(&g, B, C, H, W, 0, kh, kw, sh, sw, ph, pw, dh, dw, 1, dg);
    for (int b = 0; b < B; ++b)
        im2col_one(&g, input + (size_t)b * C * H * W, offset + b * off_bstride, mask ? mask + b * mask_bstride : NULL,
                   columns + (size_t)b * C * kh * kw * g.Ho * g.Wo);
}

/* deform_conv_cuda.cpp:486-564.  weight [Cout, C/group, kh, kw]; output [B, Cout, Ho, Wo]. */
void FN(dcn_forward)(const REAL *input, const REAL *weight, const REAL *bias, const REAL *offset, int64_t off_bstride,
                     const REAL *mask, int64_t mask_bstride, int B, int C, int H, int W, int Cout, int kh, int kw,
                     int sh, int sw, int ph, int pw, int dh, int dw, int group, int dg, int with_bias, REAL *output) {
    DcnGeo g; geo_fill(&g, B, C, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, group, dg);
    const int K = kh * kw, P = g.Ho * g.Wo, Cg = C / group, Og = Cout / group;
    REAL *col = (REAL *)malloc(sizeof(REAL) * (size_t)C * K * P);
    for (int b = 0; b < B; ++b) {
        im2col_one(&g, input + (size_t)b * C * H * W, offset + b * off_bstride, mask ? mask + b * mask_bstride : NULL, col);
        for (int gr = 0; gr < group; ++gr)
            for (int o = 0; o < Og; ++o) {
                REAL *out = output + (((size_t)b * Cout) + gr * Og + o) * P;
                for (int p = 0; p < P; ++p) out[p] = 0;
                const REAL *wrow = weight + (size_t)(gr * Og + o) * Cg * K;
                for (int r = 0; r < Cg * K; ++r) {
                    const REAL wv = wrow[r];
                    const REAL *crow = col + ((size_t)gr * Cg * K + r) * P;
                    for (int p = 0; p < P; ++p) out[p] += wv * crow[p];
                }
                if (with_bias) for (int p = 0; p < P; ++p) out[p] += bias[gr * Og + o];
            }
    }
    free(col);
}

/* deform_conv_cuda.cpp:566-679.  grad_* are ACCUMULATED INTO (caller zero-fills, functions/deform_conv.py:150-154),
 * except grad_offset / grad_mask entries, which K10 assigns. */
void FN(dcn_backward)(const REAL *input, const REAL *weight, const REAL *offset, int64_t off_bstride, const REAL *mask,
                      int64_t mask_bstride, const REAL *grad_output, int B, int C, int H, int W, int Cout, int kh, int kw,
                      int sh, int sw, int ph, int pw, int dh, int dw, int group, int dg, int with_bias,
                      REAL *grad_input, REAL *grad_weight, REAL *grad_bias, REAL *grad_offset, int64_t goff_bstride,
                      REAL *grad_mask, int64_t gmask_bstride) {
    DcnGeo g; geo_fill(&g, B, C, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, group, dg);
    const int K = kh * kw, P = g.Ho * g.Wo, Cg = C / group, Og = Cout / group;
    REAL *col = (REAL *)malloc(sizeof(REAL) * (size_t)C * K * P);
    for (int b = 0; b < B; ++b) {
        const REAL *im = input + (size_t)b * C * H * W;
        const REAL *off = offset + b * off_bstride;
        const REAL *msk = mask ? mask + b * mask_bstride : NULL;
        const REAL *go = grad_output + (size_t)b * Cout * P;
        /* columns = W^T . grad_output (:611-614) */
        for (int gr = 0; gr < group; ++gr)
            for (int r = 0; r < Cg * K; ++r) {
                REAL *crow = col + ((size_t)gr * Cg * K + r) * P;
                for (int p = 0; p < P; ++p) crow[p] = 0;
                for (int o = 0; o < Og; ++o) {
                    const REAL wv = weight[(size_t)(gr * Og + o) * Cg * K + r];
                    const REAL *gor = go + (size_t)(gr * Og + o) * P;
                    for (int p = 0; p < P; ++p) crow[p] += wv * gor[p];
                }
            }
        col2im_coord_one(&g, col, im, off, msk, grad_offset + b * goff_bstride, grad_mask ? grad_mask + b * gmask_bstride : NULL);
        col2im_one(&g, col, off, msk, grad_input + (size_t)b * C * H * W);
        im2col_one(&g, im, off, msk, col);
        for (int gr = 0; gr < group; ++gr)
            for (int o = 0; o < Og; ++o) {
                const REAL *gor = go + (size_t)(gr * Og + o) * P;
                REAL *gw = grad_weight + (size_t)(gr * Og + o) * Cg * K;
                for (int r = 0; r < Cg * K; ++r) {
                    const REAL *crow = col + ((size_t)gr * Cg * K + r) * P;
                    REAL acc = 0;
                    for (int p = 0; p < P; ++p) acc += gor[p] * crow[p];
                    gw[r] += acc;
                }
                if (with_bias) {
                    REAL acc = 0;
                    for (int p = 0; p < P; ++p) acc += gor[p];
                    grad_bias[gr * Og + o] += acc;
                }
            }
    }
    free(col);
}
