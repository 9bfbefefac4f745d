// Persistent single-sequence LSTM layer (sm_100a): the recurrence of the (Bi)RNN baseline when ONE long stream is
// evaluated (reference scripts/evaluate_real.py:39-61 feeds the hold-out subject's ~15k frames as a batch of one).
//
// A step of a batch-1 LSTM is a matrix-VECTOR product: run as a GEMM launch per step it streams all of W_hh and W_ih
// out of L2 for one useful row of a 128-row tile (measured: 35 us per step).  Here
//   * the input projection  XW[t] = W_ih x_t + b_ih + b_hh  of ALL time steps is one ordinary GEMM (M = F rows) on the
//     tensor-core job executor beforehand, and
//   * this kernel keeps W_hh resident in SHARED MEMORY for the whole sequence: every CTA owns U hidden units of one
//     direction (their 4U gate rows of W_hh as fp16, <= 128 KB), all CTAs are co-resident (cooperative launch) and
//     exchange the new hidden vector through a double-buffered fp32 array in L2 guarded by one monotonic counter per
//     direction (release / acquire).  Cell and hidden state stay fp32; only W_hh is rounded (to fp16).
// Packed-sequence semantics as in gemm_jobs.h: for t >= len the state is carried and the output row is ignored
// downstream (consumers mask it).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/empose_b200.h"
#include "common.cuh"
#include "rnn_persistent.h"

namespace empose {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) lstm_persistent_kernel(LstmPersistentParams p) {
    extern __shared__ __align__(16) uint8_t smem_lp[];
    const int H = p.H, U = p.U, C = p.C;
    const int rows = 4 * U;                          // gate rows owned by this CTA: row = gate * U + unit
    const int pitch = H * 2 + 64;                    // bytes per W row in shared memory (bank shift between rows)
    uint8_t* w_s = smem_lp;
    float* h_s = reinterpret_cast<float*>(smem_lp + (size_t)rows * pitch);      // [H] previous hidden vector
    float* pre_s = h_s + H;                                                      // [4U] gate pre-activations
    float* c_s = pre_s + rows;                                                   // [U] cell state
    const int dir = blockIdx.x / C, cta = blockIdx.x % C, u0 = cta * U;
    const int tid = threadIdx.x;
    const __half* W = p.w_hh[dir];
    // ---- W_hh slice -> shared memory (once) ----
    const int chunks_per_row = H / 8;                // 16-byte chunks
    for (int i = tid; i < rows * chunks_per_row; i += kThreads) {
        const int r = i / chunks_per_row, ck = i % chunks_per_row;
        const int g = r / U, u = r % U;
        const uint4 v = *reinterpret_cast<const uint4*>(W + ((size_t)(g * H + u0 + u)) * H + ck * 8);
        *reinterpret_cast<uint4*>(w_s + (size_t)r * pitch + ck * 16) = v;
    }
    if (tid < U) c_s[tid] = p.c0 ? p.c0[(size_t)dir * H + u0 + tid] : 0.0f;
    __syncthreads();

    const int tpr = kThreads / rows;                 // threads per gate row (power of two, <= 32)
    const int row = tid / tpr, part = tid % tpr;
    const float* xw = p.xw[dir];
    float* hx = p.hx + (size_t)dir * 2 * H;
    unsigned* counter = p.counters + dir * 32;       // one counter per direction, 128 bytes apart
    const uint8_t* wrow = w_s + (size_t)row * pitch;
    const int gate_col = (row / U) * H + u0 + (row % U);   // column of this row's pre-activation in XW (torch gate order)

    for (int s = 0; s < p.F; ++s) {
        const int t = dir == 0 ? s : p.F - 1 - s;
        const bool live = t < p.len;
        // the input projection of this step does not depend on the recurrence: fetch it before waiting
        float xw_val = 0.0f;
        if (part == 0 && live) xw_val = __ldg(xw + (size_t)t * 4 * H + gate_col);
        if (s > 0) {
            if (tid == 0) {
                const unsigned need = (unsigned)C * (unsigned)s;
                unsigned spins = 0;
                while (ld_acquire(counter) < need) {
                    if (++spins > (1u << 22)) __trap();          // a protocol bug must surface as an error, not as a hung GPU
                }
            }
            __syncthreads();
        }
        // previous hidden vector of my direction (written by all CTAs of the direction): bypass L1
        const float* hprev = s == 0 ? (p.h0 ? p.h0 + (size_t)dir * H : nullptr) : hx + (size_t)((s - 1) & 1) * H;
        for (int k = tid * 4; k < H; k += kThreads * 4) {
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (hprev) v = __ldcg(reinterpret_cast<const float4*>(hprev + k));
            *reinterpret_cast<float4*>(h_s + k) = v;
        }
        __syncthreads();
        if (live) {
            float acc = 0.0f;
            for (int ck = part; ck < chunks_per_row; ck += tpr) {
                const uint4 wv = *reinterpret_cast<const uint4*>(wrow + ck * 16);
                const float4 ha = *reinterpret_cast<const float4*>(h_s + ck * 8);
                const float4 hb = *reinterpret_cast<const float4*>(h_s + ck * 8 + 4);
                const float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&wv.x));
                const float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&wv.y));
                const float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(&wv.z));
                const float2 w3 = __half22float2(*reinterpret_cast<const __half2*>(&wv.w));
                acc = fmaf(w0.x, ha.x, acc); acc = fmaf(w0.y, ha.y, acc); acc = fmaf(w1.x, ha.z, acc); acc = fmaf(w1.y, ha.w, acc);
                acc = fmaf(w2.x, hb.x, acc); acc = fmaf(w2.y, hb.y, acc); acc = fmaf(w3.x, hb.z, acc); acc = fmaf(w3.y, hb.w, acc);
            }
            for (int off = tpr >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (part == 0) pre_s[row] = acc + xw_val;
        }
        __syncthreads();
        if (tid < U) {
            const int u = u0 + tid;
            float h = h_s[u];                                  // carried when the step is padded
            if (live) {
                const float gi = pre_s[tid], gf = pre_s[U + tid], gg = pre_s[2 * U + tid], go = pre_s[3 * U + tid];
                const float c = sigmoid_f(gf) * c_s[tid] + sigmoid_f(gi) * tanh_f(gg);
                c_s[tid] = c;
                h = sigmoid_f(go) * tanh_f(c);
            }
            hx[(size_t)(s & 1) * H + u] = h;
            store_operand(p.hseq, (size_t)t * p.hseq_pitch + (size_t)dir * H + u, h, p.hseq_mode);
            if (s == p.F - 1) {
                if (p.h_out) p.h_out[(size_t)dir * H + u] = h;
                if (p.c_out) p.c_out[(size_t)dir * H + u] = c_s[tid];
            }
        }
        __syncthreads();                                       // all h of this CTA are written ...
        if (tid == 0) {
            __threadfence();
            red_release_add(counter, 1u);                      // ... and published with release semantics
        }
    }
}

}  // namespace

size_t lstm_persistent_smem_bytes(int H, int U) {
    return (size_t)4 * U * (H * 2 + 64) + (size_t)(H + 4 * U + U) * sizeof(float);
}

bool lstm_persistent_pick(int H, int dirs, int num_sms, int* C_out, int* U_out) {
    if (H % 8) return false;
    for (int C = num_sms / dirs; C >= 1; --C) {
        if (H % C) continue;
        const int U = H / C;
        const int rows = 4 * U;
        if (rows > kThreads || kThreads % rows) continue;           // whole threads per gate row
        if (kThreads / rows > 32) continue;                          // ... reduced inside one warp
        if (lstm_persistent_smem_bytes(H, U) > 200 * 1024) continue;
        *C_out = C; *U_out = U;
        return true;
    }
    return false;
}

int launch_lstm_persistent(const LstmPersistentParams& p, cudaStream_t s) {
    const size_t smem = lstm_persistent_smem_bytes(p.H, p.U);
    static size_t configured = 0;
    if (smem > configured) {
        EMPOSE_CUDA_TRY(cudaFuncSetAttribute(lstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    EMPOSE_CUDA_TRY(cudaMemsetAsync(p.counters, 0, 2 * 32 * sizeof(unsigned), s));
    LstmPersistentParams params = p;
    void* args[] = {&params};
    // cooperative launch: the grid-wide hand-off needs every CTA resident, and the runtime refuses the launch otherwise
    EMPOSE_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)lstm_persistent_kernel, dim3(p.C * p.dirs), dim3(kThreads), args, smem, s));
    return EMPOSE_OK;
}

}  // namespace empose
