// Shared device/host utilities for the empose_b200 CUDA sources.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace empose {

// ---- error plumbing (host) --------------------------------------------------------------------
void set_last_error(const std::string& msg);

#define EMPOSE_CUDA_TRY(expr)                                                                         \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            ::empose::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +      \
                                     " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");         \
            return EMPOSE_E_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

// ---- numeric helpers (device) -----------------------------------------------------------------
// Round-to-nearest fp32 -> tf32 (10-bit mantissa), kept in an fp32 container.  tcgen05 kind::tf32
// ignores the low 13 mantissa bits of its operands, so values are rounded where they are PRODUCED
// (weights at pack time, activations in the epilogue that writes them) to get round-to-nearest
// instead of truncation.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
    // tanh(x) = 1 - 2/(exp(2x)+1); accurate to a few ulp with the fast exp, saturates cleanly
    float e = __expf(2.0f * x);
    return 1.0f - 2.0f / (e + 1.0f);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace empose
