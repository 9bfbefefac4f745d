// fp32 FFMA job executor: the exact-arithmetic mode (EMPOSE_PRECISION_FP32) of the GEMM jobs.
// Same job descriptors and epilogues as the tcgen05 executor; one launch per job, one CTA per
// (128-row tile, 32-column chunk), one thread per row.  Built for parity studies, not for speed.
#include "../../include/empose_b200.h"
#include "gemm_tc.h"

namespace empose {

namespace {

__global__ void __launch_bounds__(kTileM) gemm_simt_kernel(const GemmJob* __restrict__ jobs, int job_index) {
    const GemmJob j = jobs[job_index];
    __shared__ float a_s[kChunkK][kTileM + 1];
    __shared__ float w_s[kChunkK][33];
    __shared__ float epi_stage[(kTileM / 32) * kStageFloats];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * kTileM;
    const int c0 = blockIdx.y * 32;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;

    for (int seg = 0; seg < 2; ++seg) {
        const int kseg = j.a_k[seg];
        for (int k0 = 0; k0 < kseg; k0 += kChunkK) {
            for (int q = 0; q < kTileM * kChunkK / kTileM; ++q) {          // 32 loads per thread
                const int i = tid + kTileM * q;
                const int r = i / kChunkK, kk = i % kChunkK;
                float v = 0.0f;
                if (m0 + r < j.m_rows && k0 + kk < kseg) v = j.a_ptr[seg][(int64_t)(m0 + r) * j.a_stride[seg] + k0 + kk];
                a_s[kk][r] = v;
            }
            for (int q = 0; q < 32 * kChunkK / kTileM; ++q) {              // 8 loads per thread
                const int i = tid + kTileM * q;
                const int c = i / kChunkK, kk = i % kChunkK;
                float v = 0.0f;
                if (c0 + c < j.n_count && k0 + kk < kseg)
                    v = j.w_ptr[(int64_t)(j.n_begin + c0 + c) * j.w_ld + j.w_koff[seg] + k0 + kk];
                w_s[kk][c] = v;
            }
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < kChunkK; ++kk) {
                const float a = a_s[kk][tid];
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c] = fmaf(a, w_s[kk][c], acc[c]);
            }
            __syncthreads();
        }
    }
    epilogue_chunk(j, m0 + (tid & ~31), tid & 31, c0, acc, epi_stage + (tid >> 5) * kStageFloats);
}

}  // namespace

int simt_launch(const GemmJob* d_jobs, const GemmJob* h_jobs, int job_begin, int job_count, int m_tiles,
                cudaStream_t stream, int64_t* launch_counter) {
    for (int i = 0; i < job_count; ++i) {
        const GemmJob& j = h_jobs[job_begin + i];
        dim3 grid(m_tiles, (j.n_count + 31) / 32);
        gemm_simt_kernel<<<grid, kTileM, 0, stream>>>(d_jobs, job_begin + i);
        if (launch_counter) ++*launch_counter;
    }
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

}  // namespace empose
