// Host-side harness: runs the PRODUCT's per-pixel routines (megreader_b200/csrc/input_core.cuh, the code the CUDA kernels in
// input_pipeline.cu execute) on the CPU, with the same flat-index decomposition, so that tests can compare them with the
// oracle without a GPU.  Built on demand by tests/test_input_core_host.py with g++ (no CUDA involved).
#include "input_core.cuh"

extern "C" {

void host_resize_normalize(const void *src, int src_is_u8, const int64_t *offsets, const int *heights, const int *widths,
                           const int *valid_w, int N, int dst_h, int dst_w, const double *mean3, float *out) {
    const int64_t total = (int64_t)N * 3 * dst_h * dst_w;
    for (int64_t idx = 0; idx < total; ++idx) {
        int n, c, y, x;
        mr_input::decode_index(idx, dst_h, dst_w, n, c, y, x);
        const int vw = valid_w ? valid_w[n] : dst_w;
        out[idx] = src_is_u8
            ? mr_input::resize_normalize_value((const unsigned char *)src + offsets[n], heights[n], widths[n], dst_h, vw, y, x, c, mean3[c])
            : mr_input::resize_normalize_value((const float *)src + offsets[n], heights[n], widths[n], dst_h, vw, y, x, c, mean3[c]);
    }
}

void host_pack_labels(const unsigned char *text, const int64_t *offsets, int N, const int *lut, int max_size, int *labels,
                      int *lengths) {
    for (int n = 0; n < N; ++n) {
        const int len = (int)(offsets[n + 1] - offsets[n]);
        for (int pos = 0; pos < max_size; ++pos) labels[(int64_t)n * max_size + pos] = mr_input::pack_label_value(text + offsets[n], len, pos, lut);
        lengths[n] = len < max_size ? len : max_size;
    }
}

}
