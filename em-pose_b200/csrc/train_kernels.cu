// Training-only CUDA kernels: see train_kernels.h.  All of this is HBM-bound elementwise / reduction work
// (coalesced along the 512-wide feature axis); the GEMM-shaped parts of the training step run on the job
// executors (gemm_tc.cu / gemm_simt.cu).
#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_math.h"
#include "gemm_jobs.h"
#include "train_kernels.h"

namespace empose {

namespace {

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }
__device__ __forceinline__ float maybe_round(float x, int round_out) { return round_out ? round_tf32(x) : x; }

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(const PackOp* __restrict__ ops) {
    const PackOp op = ops[blockIdx.y];
    const int64_t total = (int64_t)op.rows * op.cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(idx / op.cols), k = (int)(idx % op.cols);
        const int sr = op.lstm_map ? lstm_gate_of_packed(n) * op.hidden + lstm_unit_of_packed(n) : n;
        const int64_t si = (int64_t)sr * op.src_rs + (int64_t)k * op.src_cs;
        float v = op.src[si];
        if (op.src2) v += op.src2[si];
        op.dst[(int64_t)n * op.dst_ld + k] = maybe_round(v, op.round);
    }
}

// ---------------------------------------------------------------------------------------------
// block (32, 8): thread (cx, ry) walks rows ry, ry + 8 * gridDim.y, ... of one segment
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ z, int64_t ld, int R, int n,
                                                        double* __restrict__ sums) {
    __shared__ double s_sum[8][33], s_sq[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int seg = blockIdx.z;
    double a = 0.0, b = 0.0;
    if (c < n) {
        const float* base = z + ((int64_t)seg * R) * ld + c;
        for (int r = blockIdx.y * 8 + threadIdx.y; r < R; r += 8 * gridDim.y) {
            const double v = (double)base[(int64_t)r * ld];
            a += v; b += v * v;
        }
    }
    s_sum[threadIdx.y][threadIdx.x] = a; s_sq[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y == 0 && c < n) {
        for (int q = 1; q < 8; ++q) { a += s_sum[q][threadIdx.x]; b += s_sq[q][threadIdx.x]; }
        atomicAdd(&sums[((int64_t)seg * 2 + 0) * n + c], a);
        atomicAdd(&sums[((int64_t)seg * 2 + 1) * n + c], b);
    }
}

// `count` = rows behind every sum: R, or the rows of ALL ranks when the sums were all-reduced (SyncBatchNorm)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int S, int n, float eps, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* running_mean, float* running_var, float momentum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * n) return;
    const int seg = i / n, c = i % n;
    const double m = sums[((int64_t)seg * 2) * n + c] / count;
    double var = sums[((int64_t)seg * 2 + 1) * n + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[i] = (float)m;
    invstd[i] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {      // torch.nn.BatchNorm1d: running_var takes the unbiased estimate
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ z, int64_t ld, int R, int n,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ alpha, int round_out, float* __restrict__ a,
                                                       int64_t a_ld, int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t row = idx / n;
    const int c = (int)(idx % n);
    const int seg = (int)(row / R);
    float y = z[row * ld + c];
    if (gamma) y = gamma[c] * ((y - mean[seg * n + c]) * invstd[seg * n + c]) + beta[c];
    const float al = alpha[0];
    y = y > 0.0f ? y : al * y;
    a[row * a_ld + c] = maybe_round(y, round_out);
}

// Same as bn_apply_kernel for n and both pitches multiples of 4 (the hidden layers): a thread owns FOUR consecutive columns of
// kApplyVecRows rows, so the per-column constants are formed once and every access is 16 bytes -- the pass is pure HBM
// traffic (read z, write a), which the scalar form reaches only a third of.
constexpr int kApplyVecRows = 8;
__global__ void __launch_bounds__(256) bn_apply_vec_kernel(const float* __restrict__ z, int64_t ld, int R, int n,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ alpha, int round_out, float* __restrict__ a,
                                                           int64_t a_ld, int S) {
    const int n4 = n >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = (int)(t % n4);
    const int64_t rb = t / n4;                         // row block
    const int64_t row0 = rb * kApplyVecRows;
    if (row0 >= (int64_t)S * R) return;
    const int c = c4 * 4;
    const float al = alpha[0];
    int seg = -1;
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f), mu = be, is = g;
    if (gamma) { g = *reinterpret_cast<const float4*>(gamma + c); be = *reinterpret_cast<const float4*>(beta + c); }
    const int64_t row_end = min(row0 + kApplyVecRows, (int64_t)S * R);
    for (int64_t row = row0; row < row_end; ++row) {
        const int sg = (int)(row / R);
        if (gamma && sg != seg) {
            seg = sg;
            mu = *reinterpret_cast<const float4*>(mean + seg * n + c);
            is = *reinterpret_cast<const float4*>(invstd + seg * n + c);
        }
        const float4 v = *reinterpret_cast<const float4*>(z + row * ld + c);
        float y[4] = {v.x, v.y, v.z, v.w};
        if (gamma) {          // the scalar kernel's association: gamma * ((z - mean) * invstd) + beta
            y[0] = g.x * ((y[0] - mu.x) * is.x) + be.x; y[1] = g.y * ((y[1] - mu.y) * is.y) + be.y;
            y[2] = g.z * ((y[2] - mu.z) * is.z) + be.z; y[3] = g.w * ((y[3] - mu.w) * is.w) + be.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { y[q] = y[q] > 0.0f ? y[q] : al * y[q]; y[q] = maybe_round(y[q], round_out); }
        *reinterpret_cast<float4*>(a + row * a_ld + c) = make_float4(y[0], y[1], y[2], y[3]);
    }
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ da, int64_t da_ld, const float* __restrict__ z,
                                                            int64_t z_ld, int R, int n, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ alpha,
                                                            double* __restrict__ sums) {
    __shared__ double sh[3][8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int seg = blockIdx.z;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (c < n) {
        const float al = alpha[0];
        const float mu = gamma ? mean[seg * n + c] : 0.0f, is = gamma ? invstd[seg * n + c] : 1.0f;
        const float g = gamma ? gamma[c] : 1.0f, be = gamma ? beta[c] : 0.0f;
        for (int r = blockIdx.y * 8 + threadIdx.y; r < R; r += 8 * gridDim.y) {
            const int64_t row = (int64_t)seg * R + r;
            const float xh = (z[row * z_ld + c] - mu) * is;
            const float y = g * xh + be;
            const float d = da[row * da_ld + c];
            const float dy = y > 0.0f ? d : al * d;
            s1 += (double)dy; s2 += (double)dy * (double)xh;
            if (!(y > 0.0f)) s3 += (double)d * (double)y;
        }
    }
    sh[0][threadIdx.y][threadIdx.x] = s1; sh[1][threadIdx.y][threadIdx.x] = s2; sh[2][threadIdx.y][threadIdx.x] = s3;
    __syncthreads();
    if (threadIdx.y == 0 && c < n) {
        for (int q = 1; q < 8; ++q) { s1 += sh[0][q][threadIdx.x]; s2 += sh[1][q][threadIdx.x]; s3 += sh[2][q][threadIdx.x]; }
        atomicAdd(&sums[((int64_t)seg * 3 + 0) * n + c], s1);
        atomicAdd(&sums[((int64_t)seg * 3 + 1) * n + c], s2);
        atomicAdd(&sums[((int64_t)seg * 3 + 2) * n + c], s3);
    }
}

// grid (rows / kApplyRows, S): a block owns kApplyRows rows of one segment; per-column constants are formed once per block
constexpr int kApplyRows = 64;
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ da, int64_t da_ld, const float* __restrict__ z,
                                                           int64_t z_ld, int R, int n, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ alpha,
                                                           const double* __restrict__ sums, double count, int round_out,
                                                           float* __restrict__ dz, int64_t dz_ld) {
    extern __shared__ float cst[];              // [6][n]: mean, invstd, gamma, beta, m1, m2
    const int seg = blockIdx.y;
    float *c_mu = cst, *c_is = cst + n, *c_g = cst + 2 * n, *c_b = cst + 3 * n, *c_m1 = cst + 4 * n, *c_m2 = cst + 5 * n;
    if (gamma)
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            c_mu[c] = mean[seg * n + c]; c_is[c] = invstd[seg * n + c]; c_g[c] = gamma[c]; c_b[c] = beta[c];
            c_m1[c] = (float)(sums[((int64_t)seg * 3 + 0) * n + c] / count);
            c_m2[c] = (float)(sums[((int64_t)seg * 3 + 1) * n + c] / count);
        }
    __syncthreads();
    const float al = alpha[0];
    const int r0 = blockIdx.x * kApplyRows;
    const int rows = min(kApplyRows, R - r0);
    if ((n & 3) == 0 && (da_ld & 3) == 0 && (z_ld & 3) == 0 && (dz_ld & 3) == 0 && (((uintptr_t)da | (uintptr_t)z | (uintptr_t)dz) & 15) == 0) {
        // 16-byte accesses: four columns per thread and step (same arithmetic, same association as the scalar loop below)
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < rows * n4; i += blockDim.x) {
            const int r = i / n4, c = (i - r * n4) * 4;
            const int64_t row = (int64_t)seg * R + r0 + r;
            const float4 d4 = *reinterpret_cast<const float4*>(da + row * da_ld + c), z4 = *reinterpret_cast<const float4*>(z + row * z_ld + c);
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (gamma) {
                    const float xh = (zz[q] - c_mu[c + q]) * c_is[c + q];
                    const float y = c_g[c + q] * xh + c_b[c + q];
                    const float dy = y > 0.0f ? dd[q] : al * dd[q];
                    o[q] = c_g[c + q] * c_is[c + q] * (dy - c_m1[c + q] - xh * c_m2[c + q]);
                } else {
                    o[q] = zz[q] > 0.0f ? dd[q] : al * dd[q];
                }
                o[q] = maybe_round(o[q], round_out);
            }
            *reinterpret_cast<float4*>(dz + row * dz_ld + c) = make_float4(o[0], o[1], o[2], o[3]);
        }
        return;
    }
    for (int i = threadIdx.x; i < rows * n; i += blockDim.x) {
        const int r = i / n, c = i - r * n;
        const int64_t row = (int64_t)seg * R + r0 + r;
        const float d = da[row * da_ld + c];
        float out;
        if (gamma) {
            const float xh = (z[row * z_ld + c] - c_mu[c]) * c_is[c];
            const float y = c_g[c] * xh + c_b[c];
            const float dy = y > 0.0f ? d : al * d;
            out = c_g[c] * c_is[c] * (dy - c_m1[c] - xh * c_m2[c]);
        } else {
            const float y = z[row * z_ld + c];
            out = y > 0.0f ? d : al * d;
        }
        dz[row * dz_ld + c] = maybe_round(out, round_out);
    }
}

__global__ void __launch_bounds__(256) bn_param_grads_kernel(const double* __restrict__ sums, int S, int n, float* g_gamma,
                                                             float* g_beta, float* g_alpha) {
    __shared__ double part[256];
    double a3 = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        double a1 = 0.0, a2 = 0.0;
        for (int s = 0; s < S; ++s) {
            a1 += sums[((int64_t)s * 3 + 0) * n + c];
            a2 += sums[((int64_t)s * 3 + 1) * n + c];
            a3 += sums[((int64_t)s * 3 + 2) * n + c];
        }
        if (g_gamma) g_gamma[c] += (float)a2;
        if (g_beta) g_beta[c] += (float)a1;
    }
    part[threadIdx.x] = a3;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) g_alpha[0] += (float)part[0];
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_sum_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int n,
                                                      double* __restrict__ scratch) {
    __shared__ double sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double a = 0.0;
    if (c < n)
        for (int64_t r = blockIdx.y * 8 + threadIdx.y; r < rows; r += 8 * gridDim.y) a += (double)src[r * ld + c];
    sh[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0 && c < n) {
        for (int q = 1; q < 8; ++q) a += sh[q][threadIdx.x];
        atomicAdd(&scratch[c], a);
    }
}
__global__ void col_sum_finish_kernel(const double* __restrict__ scratch, int n, float* dst, float* dst2) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const float v = (float)scratch[c];
    dst[c] += v;
    if (dst2) dst2[c] += v;
}

__global__ void __launch_bounds__(256) dw_reduce_kernel(const DwReduce* __restrict__ table) {
    const DwReduce d = table[blockIdx.y];
    const int n4 = d.count >> 2;                          // counts are multiples of 4 (row lengths are), buffers 16-byte aligned
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 acc = reinterpret_cast<float4*>(d.grad)[i];
        for (int q = 0; q < d.splits; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(d.partial + (size_t)q * d.count) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(d.grad)[i] = acc;
    }
    if (blockIdx.x == 0)
        for (int i = (n4 << 2) + threadIdx.x; i < d.count; i += blockDim.x) {
            float acc = d.grad[i];
            for (int q = 0; q < d.splits; ++q) acc += d.partial[(size_t)q * d.count + i];
            d.grad[i] = acc;
        }
}

// 64 x 64 tile transpose through shared memory; block (32, 8): sixteen loads per thread in flight before the first store
// (with 32 x 32 tiles and four it ran at half the HBM rate: 88 us for a [65536 x 512] operand)
constexpr int kTrTile = 64;
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, int64_t src_ld, int64_t rows, int n, int shift,
                                                        int F, int round_out, float* __restrict__ dst, int64_t dst_ld) {
    __shared__ float tile[kTrTile][kTrTile + 1];
    const int64_t r0 = (int64_t)blockIdx.x * kTrTile;
    const int c0 = blockIdx.y * kTrTile;
    float v[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int q = threadIdx.y + 8 * k;
            const int64_t r = r0 + q;
            const int c = c0 + threadIdx.x + 32 * h;
            float x = 0.0f;
            if (r < rows && c < n) {
                if (!shift) x = src[r * src_ld + c];
                else if (r % F != 0) x = src[(r - 1) * src_ld + c];
            }
            v[h][k] = x;
        }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) tile[threadIdx.y + 8 * k][threadIdx.x + 32 * h] = v[h][k];
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int q = threadIdx.y + 8 * k;
            const int c = c0 + q;
            const int64_t r = r0 + threadIdx.x + 32 * h;
            if (c < n && r < rows) dst[(int64_t)c * dst_ld + r] = maybe_round(tile[threadIdx.x + 32 * h][q], round_out);
        }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(LstmCellBwdParams p) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)p.B * p.H) return;
    const int b = (int)(idx / p.H), u = (int)(idx % p.H);
    const int64_t row = (int64_t)b * p.F + p.t;
    float* dg = p.dgates + row * 4 * p.H + u;
    if (p.t >= p.seq_len[b]) {        // padded step: the state was carried, the output is zero (layers.py:146-153)
        dg[0] = 0.0f; dg[p.H] = 0.0f; dg[2 * p.H] = 0.0f; dg[3 * p.H] = 0.0f;
        return;
    }
    // the forward epilogue keeps the activated gates in ITS column order: 32-column groups of [i f g o] x 8 units
    const float* g = p.gates + row * 4 * p.H + (u >> 3) * 32 + (u & 7);
    const float gi = g[0], gf = g[8], gg = g[16], go = g[24];
    const float c = p.c_seq[row * p.H + u];
    const float c_prev = p.t > 0 ? p.c_seq[(row - 1) * p.H + u] : 0.0f;
    float dh = p.dh_out[row * p.H + u];
    float dc = 0.0f;
    if (!p.last) { dh += p.dh_rec[idx]; dc = p.dc_rec[idx]; }
    const float tc = tanhf(c);
    dc += dh * go * (1.0f - tc * tc);
    p.dc_rec[idx] = dc * gf;
    dg[0] = maybe_round(dc * gg * gi * (1.0f - gi), p.round_out);
    dg[p.H] = maybe_round(dc * c_prev * gf * (1.0f - gf), p.round_out);
    dg[2 * p.H] = maybe_round(dc * gi * (1.0f - gg * gg), p.round_out);
    dg[3 * p.H] = maybe_round(dh * tc * go * (1.0f - go), p.round_out);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sign_f(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// per-iterate total gradient wrt theta_i (c < 66) or beta_i (c >= 66) of frame (b, f)
__device__ __forceinline__ float seed_term(const SeedParams& p, int i, int b, int f, int c, float m_over_blen) {
    const int64_t R = (int64_t)p.B * p.F, row = (int64_t)b * p.F + f;
    const float w = (i < p.N) ? (p.recon_w / (float)(p.N + 1) + (p.side_effect ? 1.0f : 0.0f)) : 1.0f;
    const float inv_bf = 1.0f / (float)R;
    if (c < kPoseDim) {
        const float th = p.pose_hist[((int64_t)i * R + row) * kPoseDim + c];
        const float gt = p.pose_gt[row * kPoseDim + c];
        return p.pose_w / (float)(p.N + 1) * m_over_blen / (float)kPoseDim * sign_f(th - gt) +
               w * p.g_theta[((int64_t)i * R + row) * kPoseDim + c] * inv_bf;
    }
    const int k = c - kPoseDim;
    const float be = p.shape_hist[((int64_t)i * R + row) * kBetas + k];
    const float gt = p.shape_gt[(int64_t)b * kBetas + k];
    return p.shape_w / (float)(p.N + 1) * m_over_blen / (float)kBetas * sign_f(be - gt) +
           w * p.g_beta[((int64_t)i * R + row) * kBetas + k] * inv_bf;
}

// one CTA per window; dynamic shared memory: F * 10 floats + 10 floats
__global__ void __launch_bounds__(256) seed_kernel(SeedParams p) {
    extern __shared__ float sm_seed[];
    float* tile = sm_seed;                       // [F][10]
    float* mean = sm_seed + p.F * kBetas;         // [10]
    const int b = blockIdx.x;
    const int len = p.seq_len[b];
    const int64_t R = (int64_t)p.B * p.F;
    // ---- theta: purely per element ----
    for (int e = threadIdx.x; e < p.F * kPoseDim; e += blockDim.x) {
        const int f = e / kPoseDim, c = e % kPoseDim;
        const int64_t row = (int64_t)b * p.F + f;
        const float mw = f < len ? 1.0f / ((float)p.B * (float)len) : 0.0f;
        float acc = 0.0f;
        for (int i = p.N; i >= 1; --i) {
            acc += seed_term(p, i, b, f, c, mw);
            p.d_dtheta[((int64_t)(i - 1) * R + row) * p.ld_t + c] = p.step * acc;
        }
        acc += seed_term(p, 0, b, f, c, mw);
        p.d_init[row * p.ld_i + c] = acc;
        if (p.d_init_masked) p.d_init_masked[row * p.ld_i + c] = f < len ? acc : 0.0f;
    }
    // ---- beta: level by level, with the adjoint of the window mean (models.py:529-535) ----
    for (int level = p.N; level >= 0; --level) {     // level k >= 1 produces d_dbeta[k-1]; level 0 produces d_init
        for (int e = threadIdx.x; e < p.F * kBetas; e += blockDim.x) {
            const int f = e / kBetas, k = e % kBetas;
            const float mw = f < len ? 1.0f / ((float)p.B * (float)len) : 0.0f;
            float acc = 0.0f;
            for (int i = p.N; i >= (level == 0 ? 0 : level); --i) acc += seed_term(p, i, b, f, kPoseDim + k, mw);
            tile[e] = acc;
        }
        __syncthreads();
        if (p.average_shape && threadIdx.x < kBetas) {
            float s = 0.0f;
            for (int f = 0; f < p.F; ++f) s += tile[f * kBetas + threadIdx.x];
            mean[threadIdx.x] = s / (float)p.F;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < p.F * kBetas; e += blockDim.x) {
            const int f = e / kBetas, k = e % kBetas;
            const int64_t row = (int64_t)b * p.F + f;
            const float v = p.average_shape ? mean[k] : tile[e];
            if (level >= 1) {
                p.d_dbeta[((int64_t)(level - 1) * R + row) * p.ld_b + k] = p.step * v;
            } else {
                p.d_init[row * p.ld_i + p.init_beta_col + k] = v;
                if (p.d_init_masked) p.d_init_masked[row * p.ld_i + p.init_beta_col + k] = f < len ? v : 0.0f;
            }
        }
        __syncthreads();
    }
}

// one thread per (iterate, frame); block-level reduction into 4 doubles
__global__ void __launch_bounds__(256) loss_kernel(LossParams p) {
    __shared__ double sh[4][256];
    const int64_t R = (int64_t)p.B * p.F;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (idx < (int64_t)(p.N + 1) * R) {
        const int i = (int)(idx / R);
        const int64_t row = idx % R;
        const int b = (int)(row / p.F), f = (int)(row % p.F);
        const int len = p.seq_len[b];
        const double mw = f < len ? 1.0 / ((double)p.B * (double)len) : 0.0;
        const double cw = (double)p.coef[row] / (double)R;
        if (mw > 0.0) {
            const float* th = p.pose_hist + ((int64_t)i * R + row) * kPoseDim;
            const float* tg = p.pose_gt + row * kPoseDim;
            float a = 0.0f;
            for (int c = 0; c < kPoseDim; ++c) a += fabsf(tg[c] - th[c]);
            v[0] = mw * (double)a / kPoseDim;
            const float* be = p.shape_hist + ((int64_t)i * R + row) * kBetas;
            const float* bg = p.shape_gt + (int64_t)b * kBetas;
            a = 0.0f;
            for (int c = 0; c < kBetas; ++c) a += fabsf(bg[c] - be[c]);
            v[1] = mw * (double)a / kBetas;
        }
        if (cw > 0.0) {
            const float* mp = p.markers_hist + ((int64_t)i * R + row) * 36;
            const float* mo = p.markers_ori_hist + ((int64_t)i * R + row) * 108;
            const float* me = p.meas + row * 144;
            float e = 0.0f;
            for (int s = 0; s < kSensors; ++s) {
                if (!p.sensor_active[s]) continue;
                if (p.use_pos) {
                    float q = 0.0f;
                    for (int d = 0; d < 3; ++d) { const float t = mp[s * 3 + d] - me[s * 12 + d]; q += t * t; }
                    e += sqrtf(q);
                }
                if (p.use_ori) {
                    float q = 0.0f;
                    for (int d = 0; d < 9; ++d) { const float t = mo[s * 9 + d] - me[s * 12 + 3 + d]; q += t * t; }
                    e += sqrtf(q);
                }
            }
            v[2] = cw * (double)e;
            if (p.joints_gt) {          // the FK term is evaluated on the FINAL joints for every iterate (models.py:657-660)
                const float* j = p.joints_final + row * kPoseDim;
                const float* jg = p.joints_gt + row * kPoseDim;
                float ej = 0.0f;
                for (int q = 0; q < kJoints; ++q) {
                    float s2 = 0.0f;
                    for (int d = 0; d < 3; ++d) { const float t = j[q * 3 + d] - jg[q * 3 + d]; s2 += t * t; }
                    ej += sqrtf(s2);
                }
                v[3] = cw * (double)ej;
            }
        }
    }
    for (int q = 0; q < 4; ++q) sh[q][threadIdx.x] = v[q];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int q = 0; q < 4; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 4) atomicAdd(&p.sums[threadIdx.x], sh[threadIdx.x][0]);
}

__global__ void round_kernel(float* x, int64_t rows, int n, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * n) return;
    float* p = x + (i / n) * ld + (i % n);
    *p = round_tf32(*p);
}

inline int row_chunks(int64_t rows) {
    int64_t c = (rows + 255) / 256;
    return (int)(c < 1 ? 1 : (c > 64 ? 64 : c));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
int launch_pack(const PackOp* d_ops, int n_ops, int max_elems, cudaStream_t s) {
    if (n_ops == 0) return EMPOSE_OK;
    unsigned gx = blocks_for(max_elems, 256);
    if (gx > 512) gx = 512;
    pack_kernel<<<dim3(gx, n_ops), 256, 0, s>>>(d_ops);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_col_stats(const float* z, int64_t ld, int R, int S, int n, double* sums, cudaStream_t s) {
    EMPOSE_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)S * 2 * n * sizeof(double), s));
    col_stats_kernel<<<dim3(blocks_for(n, 32), row_chunks(R), S), dim3(32, 8), 0, s>>>(z, ld, R, n, sums);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_bn_finalize(const double* sums, int64_t count, int S, int n, float eps, float* mean, float* invstd, float* running_mean,
                       float* running_var, float momentum, cudaStream_t s) {
    bn_finalize_kernel<<<blocks_for((int64_t)S * n, 256), 256, 0, s>>>(sums, (double)count, S, n, eps, mean, invstd, running_mean, running_var,
                                                                      momentum);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_bn_apply(const float* z, int64_t ld, int R, int S, int n, const float* mean, const float* invstd, const float* gamma,
                    const float* beta, const float* alpha, int round_out, float* a, int64_t a_ld, cudaStream_t s) {
    const int64_t total = (int64_t)S * R * n;
    if ((n & 3) == 0 && (ld & 3) == 0 && (a_ld & 3) == 0 && ((uintptr_t)z & 15) == 0 && ((uintptr_t)a & 15) == 0) {
        const int64_t threads = (((int64_t)S * R + kApplyVecRows - 1) / kApplyVecRows) * (n >> 2);
        bn_apply_vec_kernel<<<blocks_for(threads, 256), 256, 0, s>>>(z, ld, R, n, mean, invstd, gamma, beta, alpha, round_out, a, a_ld, S);
        EMPOSE_CUDA_TRY(cudaGetLastError());
        return EMPOSE_OK;
    }
    bn_apply_kernel<<<blocks_for(total, 256), 256, 0, s>>>(z, ld, R, n, mean, invstd, gamma, beta, alpha, round_out, a, a_ld, total);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_bn_bwd_reduce(const float* da, int64_t da_ld, const float* z, int64_t z_ld, int R, int S, int n, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, const float* alpha, double* sums,
                         cudaStream_t s) {
    EMPOSE_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)S * 3 * n * sizeof(double), s));
    bn_bwd_reduce_kernel<<<dim3(blocks_for(n, 32), row_chunks(R), S), dim3(32, 8), 0, s>>>(da, da_ld, z, z_ld, R, n, mean, invstd,
                                                                                           gamma, beta, alpha, sums);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_bn_bwd_apply(const float* da, int64_t da_ld, const float* z, int64_t z_ld, int R, int S, int n, const float* mean,
                        const float* invstd, const float* gamma, const float* beta, const float* alpha, const double* sums,
                        int64_t count, int round_out, float* dz, int64_t dz_ld, cudaStream_t s) {
    bn_bwd_apply_kernel<<<dim3(blocks_for(R, kApplyRows), S), 256, (size_t)6 * n * sizeof(float), s>>>(
        da, da_ld, z, z_ld, R, n, mean, invstd, gamma, beta, alpha, sums, (double)count, round_out, dz, dz_ld);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_bn_param_grads(const double* sums, int S, int n, float* g_gamma, float* g_beta, float* g_alpha, cudaStream_t s) {
    bn_param_grads_kernel<<<1, 256, 0, s>>>(sums, S, n, g_gamma, g_beta, g_alpha);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_col_sum(const float* src, int64_t ld, int64_t rows, int n, double* scratch, float* dst, float* dst2, cudaStream_t s) {
    EMPOSE_CUDA_TRY(cudaMemsetAsync(scratch, 0, (size_t)n * sizeof(double), s));
    col_sum_kernel<<<dim3(blocks_for(n, 32), row_chunks(rows)), dim3(32, 8), 0, s>>>(src, ld, rows, n, scratch);
    col_sum_finish_kernel<<<blocks_for(n, 256), 256, 0, s>>>(scratch, n, dst, dst2);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_dw_reduce(const DwReduce* d_table, int n_entries, int max_count, cudaStream_t s) {
    if (n_entries == 0) return EMPOSE_OK;
    const unsigned bx = (unsigned)std::max(1, std::min(64, (max_count / 4 + 255) / 256));
    dw_reduce_kernel<<<dim3(bx, (unsigned)n_entries), 256, 0, s>>>(d_table);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_transpose(const float* src, int64_t src_ld, int64_t rows, int n, int shift, int F, int round_out, float* dst,
                     int64_t dst_ld, cudaStream_t s) {
    transpose_kernel<<<dim3(blocks_for(rows, kTrTile), blocks_for(n, kTrTile)), dim3(32, 8), 0, s>>>(src, src_ld, rows, n, shift, F, round_out,
                                                                                           dst, dst_ld);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_lstm_cell_bwd(const LstmCellBwdParams& p, cudaStream_t s) {
    lstm_cell_bwd_kernel<<<blocks_for((int64_t)p.B * p.H, 256), 256, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_seeds(const SeedParams& p, cudaStream_t s) {
    const size_t smem = ((size_t)p.F * kBetas + kBetas) * sizeof(float);
    if (smem > 48 * 1024) { set_last_error("training windows longer than 1200 frames are not supported"); return EMPOSE_E_ARG; }
    seed_kernel<<<p.B, 256, smem, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_losses(const LossParams& p, cudaStream_t s) {
    EMPOSE_CUDA_TRY(cudaMemsetAsync(p.sums, 0, 4 * sizeof(double), s));
    loss_kernel<<<blocks_for((int64_t)(p.N + 1) * p.B * p.F, 256), 256, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

__global__ void __launch_bounds__(256) add_inplace_kernel(float4* __restrict__ dst, const float4* __restrict__ src, int64_t n4, int round_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = dst[i];
    const float4 b = __ldg(src + i);
    a.x = maybe_round(a.x + b.x, round_out); a.y = maybe_round(a.y + b.y, round_out);
    a.z = maybe_round(a.z + b.z, round_out); a.w = maybe_round(a.w + b.w, round_out);
    dst[i] = a;
}
int launch_add_inplace(float* dst, const float* src, int64_t n, int round_out, cudaStream_t s) {
    add_inplace_kernel<<<blocks_for(n / 4, 256), 256, 0, s>>>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n / 4, round_out);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_round_inplace(float* x, int64_t rows, int n, int64_t ld, cudaStream_t s) {
    round_kernel<<<blocks_for(rows * n, 256), 256, 0, s>>>(x, rows, n, ld);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

}  // namespace empose
