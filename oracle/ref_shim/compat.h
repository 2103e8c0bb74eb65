/* Force-included (-include) when oracle/build_ref.py compiles the UNMODIFIED reference native ops against the PyTorch
 * of this image.  It only restores three spellings that PyTorch removed after 1.2; no reference code is changed:
 *   AT_CHECK(cond, ...)                      -> TORCH_CHECK            (ctc2d_cuda.cu:35-42, deform_conv_cuda.cpp:65-575)
 *   AT_DISPATCH_*(tensor.type(), ...)        -> needs ::detail::scalar_type(const DeprecatedTypeProperties&)
 * TEST INFRASTRUCTURE: the product never includes this file. */
#pragma once
#ifdef __cplusplus
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
#include <c10/util/Exception.h>
#ifndef AT_CHECK
#define AT_CHECK TORCH_CHECK
#endif
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail
#endif
