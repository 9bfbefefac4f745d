// Shared device/host utilities for the empose_b200 CUDA sources.
#pragma once

#include <cuda_fp16.h>
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

// MUFU-based activations: ex2.approx + rcp.approx (2 ulp each; __fdividef returns 0 for a denominator beyond 2^126,
// which is the limit wanted here).  An IEEE division instead costs ~10 instructions and a slow-path call per use --
// the LSTM epilogue has five of these per hidden unit and row.
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
    // tanh(x) = 1 - 2/(exp(2x)+1): saturates cleanly at both ends
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// How a value that feeds a later GEMM is stored: exact fp32 (FFMA executor), tf32-rounded fp32 (kind::tf32 operands)
// or fp16 (kind::f16 operands; the buffer then holds __half elements and strides count elements).
enum OperandMode : int { OPERAND_F32 = 0, OPERAND_TF32 = 1, OPERAND_F16 = 2 };

#if defined(__CUDACC__)
__device__ __forceinline__ void store_operand(float* base, int64_t idx, float v, int mode) {
    if (mode == OPERAND_F16) reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
    else base[idx] = (mode == OPERAND_TF32) ? round_tf32(v) : v;
}
__device__ __forceinline__ float load_operand(const float* base, int64_t idx, int mode) {
    return mode == OPERAND_F16 ? __half2float(reinterpret_cast<const __half*>(base)[idx]) : base[idx];
}
#endif
inline size_t operand_bytes(int mode) { return mode == OPERAND_F16 ? 2 : 4; }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace empose
