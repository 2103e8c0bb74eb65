/* Shim for the include line `#include <THC/THC.h>` of the reference's 2D-CTC op (ops/ctc_2d/csrc/cuda/ctc2d_cuda.cu:5).
 * The header no longer ships with PyTorch; the reference uses nothing from it.  TEST INFRASTRUCTURE (oracle/_ref build). */
#pragma once
