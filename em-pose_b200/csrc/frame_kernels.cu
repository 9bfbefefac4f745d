// Per-frame CUDA kernels of the LGD loop: input assembly, estimate update + feature rows of the blend GEMM, the GENERAL
// SMPL sub-model forward / reverse kernel (per-frame state in shared memory, every phase of frame_math.h flattened over
// (frame, item) across the CTA; the production kernel for fan-form sub-models is fan_kernel.cu) and the gradient-feature
// finish.  Reference call sites are cited in frame_kernels.h and frame_math.h.
#include <stdlib.h>

#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_kernels.h"

namespace empose {

namespace {

__device__ __forceinline__ float maybe_round(float x, int round_out) { return round_out ? round_tf32(x) : x; }

// Column `col` of a feature row of the blend GEMM (`row` points at the row's first element), in the operand mode of
// the GEMM: plain fp32; tf32 value | tf32 residual (3xTF32); fp16 value | fp16 (residual * 2^11) (3xFP16).  Both
// splits carry 22 mantissa bits; the second half of the row starts kPoseFeatPad elements in.
__device__ __forceinline__ void write_feature(float v, float* row, int col, int mode) {
    if (mode == OPERAND_F16) {
        __half* h = reinterpret_cast<__half*>(row);
        const __half hi = __float2half_rn(v);
        h[col] = hi;
        h[kPoseFeatPad + col] = __float2half_rn((v - __half2float(hi)) * kSplitLoScale);
    } else if (mode == OPERAND_TF32) {
        const float hi = round_tf32(v);
        row[col] = hi;
        row[kPoseFeatPad + col] = round_tf32(v - hi);
    } else {
        row[col] = v;
    }
}
// pose features of one joint: vec(R - I) into columns [col0, col0 + 9)
__device__ __forceinline__ void write_pose_features(const float* r, float* row, int col0, int mode) {
    float R[9];
    rodrigues_fwd(r, R);
#pragma unroll
    for (int e = 0; e < 9; ++e) write_feature(R[e] - ((e % 4 == 0) ? 1.0f : 0.0f), row, col0 + e, mode);
}
// value headed for the transposed blend GEMM: dE/dvp or dE/dJ in the GEMM's operand mode
__device__ __forceinline__ float grad_operand_f32(float x, int mode) { return mode == OPERAND_TF32 ? round_tf32(x) : x; }

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prepare_kernel(PrepareParams p) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)p.R * 144) return;
    const int row = (int)(idx / 144), c = (int)(idx % 144);
    float v;
    int dst = -1;
    if (c < 36) {
        v = p.marker_pos[(int64_t)row * 36 + c];
        const int slot = p.slot_of_sensor[c / 3];
        if (p.use_pos && slot >= 0) dst = slot * 3 + c % 3;
        p.meas[(int64_t)row * 144 + (c / 3) * 12 + c % 3] = v;
    } else {
        const int e = c - 36;
        v = p.marker_oris[(int64_t)row * 108 + e];
        const int slot = p.slot_of_sensor[e / 9];
        if (p.use_ori && slot >= 0) dst = p.n_pos + slot * 9 + e % 9;
        p.meas[(int64_t)row * 144 + (e / 9) * 12 + 3 + e % 9] = v;
    }
    if (dst >= 0) {
        if (p.xin) store_operand(p.xin, (int64_t)row * p.in_stride + dst, v, p.operand_mode);
        if (p.xiter) store_operand(p.xiter, (int64_t)row * p.iter_stride + dst, v, p.operand_mode);
    }
    if (c == 0) {
        const int b = row / p.F, f = row % p.F;
        const int len = p.seq_len[b];
        float w = (f < len) ? (float)p.F / (float)len : 0.0f;
        if (p.masks) {
            const float* mk = p.masks + (int64_t)row * kSensors;
            bool all_present = true;
            for (int s = 0; s < kSensors; ++s) all_present = all_present && (mk[s] != 0.0f);
            if (!all_present) w = 0.0f;
        }
        p.coef[row] = w;
    }
}

// The same for fp16 operand buffers, kPrepRows rows per CTA: inputs are read as flat contiguous blocks, permuted in shared
// memory and written out as whole 16-byte vectors (`meas` rows and the input columns of `xin` / `xiter` rows), instead of
// one scattered 2-byte store per element (105 us for 131072 rows, three times what its bytes need).
constexpr int kPrepRows = 16;
__global__ void __launch_bounds__(256) prepare_rows_kernel(PrepareParams p) {
    __shared__ __align__(16) float sm_meas[kPrepRows * 144];
    __shared__ __align__(16) __half sm_x[kPrepRows * 144];
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * kPrepRows;
    const int nr = (int)min((int64_t)kPrepRows, (int64_t)p.R - row0);
    const int n_in = p.in_size;
    for (int i = tid; i < nr * 36; i += 256) {
        const int r = i / 36, c = i - r * 36;
        const float v = __ldg(p.marker_pos + row0 * 36 + i);
        sm_meas[r * 144 + (c / 3) * 12 + c % 3] = v;
        const int slot = p.slot_of_sensor[c / 3];
        if (p.use_pos && slot >= 0) sm_x[r * n_in + slot * 3 + c % 3] = __float2half_rn(v);
    }
    for (int i = tid; i < nr * 108; i += 256) {
        const int r = i / 108, e = i - r * 108;
        const float v = __ldg(p.marker_oris + row0 * 108 + i);
        sm_meas[r * 144 + (e / 9) * 12 + 3 + e % 9] = v;
        const int slot = p.slot_of_sensor[e / 9];
        if (p.use_ori && slot >= 0) sm_x[r * n_in + p.n_pos + slot * 9 + e % 9] = __float2half_rn(v);
    }
    if (tid < nr) {
        const int64_t row = row0 + tid;
        const int b = (int)(row / p.F), f = (int)(row % p.F);
        const int len = p.seq_len[b];
        float w = (f < len) ? (float)p.F / (float)len : 0.0f;
        if (p.masks) {
            const float* mk = p.masks + row * kSensors;
            bool all_present = true;
            for (int s = 0; s < kSensors; ++s) all_present = all_present && (mk[s] != 0.0f);
            if (!all_present) w = 0.0f;
        }
        p.coef[row] = w;
    }
    __syncthreads();
    {
        float4* dst = reinterpret_cast<float4*>(p.meas + row0 * 144);
        const float4* src = reinterpret_cast<const float4*>(sm_meas);
        for (int i = tid; i < nr * 36; i += 256) dst[i] = src[i];
    }
    const int vec = n_in >> 3;                       // 16-byte vectors per row (launch_prepare checks n_in % 8 == 0 and the pitches)
    for (int i = tid; i < nr * vec; i += 256) {
        const int r = i / vec, q = i - r * vec;
        const uint4 v = reinterpret_cast<const uint4*>(sm_x + r * n_in)[q];
        if (p.xin) reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.xin) + (row0 + r) * p.in_stride)[q] = v;
        if (p.xiter) reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.xiter) + (row0 + r) * p.iter_stride)[q] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// One CTA per window.  Frames are processed in groups of kUpdateGroup: theta of the group is kept in shared memory for
// the pose features, which are assembled in a shared-memory tile and written out as one contiguous, 16-byte-vectorised
// block (a group's rows of `pf` are adjacent in memory) instead of 18 scattered words per joint.
constexpr int kUpdateGroup = 16;
constexpr int kUpdateThreads = 256;

__global__ void __launch_bounds__(kUpdateThreads) update_kernel(UpdateParams p) {
    __shared__ float part_db[8][kBetas];
    __shared__ float mean_db[kBetas];
    __shared__ float th[kUpdateGroup][kPoseDim];
    __shared__ float be[kUpdateGroup][kBetas];
    extern __shared__ __align__(16) float tile[];          // kUpdateGroup feature rows in the operand mode's width (launch_update)
    const int b = blockIdx.x;
    const int64_t row0 = (int64_t)b * p.F;
    const int tid = threadIdx.x;
    if (p.average_shape) {
        if (tid < 8 * kBetas) {
            const int k = tid % kBetas, part = tid / kBetas;
            float acc = 0.0f;
            for (int f = part; f < p.F; f += 8) acc += p.dbeta[(row0 + f) * kBetas + k];
            part_db[part][k] = acc;
        }
        __syncthreads();
        if (tid < kBetas) {
            float acc = 0.0f;
            for (int q = 0; q < 8; ++q) acc += part_db[q][tid];
            mean_db[tid] = acc / (float)p.F;
        }
        __syncthreads();
    }
    // pad columns of the pose-feature rows (189..191 of each half) stay zero
    const int tile_words = kUpdateGroup * (p.pf_split == OPERAND_F16 ? p.pf_stride / 2 : p.pf_stride);
    for (int i = tid; i < tile_words; i += kUpdateThreads) tile[i] = 0.0f;
    for (int f0 = 0; f0 < p.F; f0 += kUpdateGroup) {
        const int nf = min(kUpdateGroup, p.F - f0);
        const int64_t g0 = row0 + f0;
        // theta rows of the group are contiguous.  Loads of a whole batch of items first, then the stores: with one item per
        // loop iteration every iteration was a full memory round trip (the compiler keeps a load behind the stores before it).
        constexpr int kBatch = (kUpdateGroup * kPoseDim + kUpdateThreads - 1) / kUpdateThreads;
        {
            float d[kBatch], t0[kBatch];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int i = tid + q * kUpdateThreads;
                d[q] = i < nf * kPoseDim ? __ldg(p.dtheta + g0 * kPoseDim + i) : 0.0f;
                t0[q] = (i < nf * kPoseDim && !p.first) ? p.theta[g0 * kPoseDim + i] : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int i = tid + q * kUpdateThreads;
                if (i < nf * kPoseDim) {
                    const int f = i / kPoseDim, c = i - f * kPoseDim;
                    const float v = p.first ? d[q] : t0[q] + p.step * d[q];
                    p.theta[g0 * kPoseDim + i] = v;
                    th[f][c] = v;
                    if (p.hist_pose) p.hist_pose[g0 * kPoseDim + i] = v;
                    if (p.xiter) store_operand(p.xiter, (g0 + f) * p.iter_stride + p.in_size + c, v, p.operand_mode);
                }
            }
        }
        for (int i = tid; i < nf * kBetas; i += kUpdateThreads) {
            const int f = i / kBetas, k = i - f * kBetas;
            const float d = p.average_shape ? mean_db[k] : p.dbeta[g0 * kBetas + i];
            const float v = p.first ? d : p.beta[g0 * kBetas + i] + p.step * d;
            p.beta[g0 * kBetas + i] = v;
            be[f][k] = v;
            if (p.hist_shape) p.hist_shape[g0 * kBetas + i] = v;
            if (p.xiter) store_operand(p.xiter, (g0 + f) * p.iter_stride + p.in_size + kPoseDim + k, v, p.operand_mode);
        }
        __syncthreads();
        // a row is pf_stride elements of 4 bytes (fp32 / tf32) or 2 bytes (fp16); rows of the group are adjacent in the tile
        const int row_words = p.pf_split == OPERAND_F16 ? p.pf_stride / 2 : p.pf_stride;
        for (int i = tid; i < nf * (kJoints - 1); i += kUpdateThreads) {
            const int f = i / (kJoints - 1), j = 1 + i - f * (kJoints - 1);
            write_pose_features(&th[f][j * 3], tile + f * row_words, (j - 1) * 9, p.pf_split);
        }
        for (int i = tid; i < nf * kBetas; i += kUpdateThreads) {            // shape columns of the feature row
            const int f = i / kBetas, k = i - f * kBetas;
            write_feature(be[f][k], tile + f * row_words, kFeatBeta + k, p.pf_split);
        }
        __syncthreads();
        float4* dst = reinterpret_cast<float4*>(p.pf + g0 * row_words);
        const float4* src = reinterpret_cast<const float4*>(tile);
        for (int i = tid; i < nf * row_words / 4; i += kUpdateThreads) dst[i] = src[i];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) pose_feature_kernel(const float* __restrict__ theta, const float* __restrict__ beta,
                                                           float* __restrict__ pf, int pf_stride, int pf_split, int R) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)R * (kJoints - 1)) return;
    const int64_t row = i / (kJoints - 1);
    const int j = 1 + (int)(i % (kJoints - 1));
    float r[3] = {theta[row * kPoseDim + j * 3], theta[row * kPoseDim + j * 3 + 1], theta[row * kPoseDim + j * 3 + 2]};
    float* frow = pf + row * (pf_split == OPERAND_F16 ? pf_stride / 2 : pf_stride);
    write_pose_features(r, frow, (j - 1) * 9, pf_split);
    if (j <= kBetas) write_feature(beta[row * kBetas + j - 1], frow, kFeatBeta + j - 1, pf_split);
}

__global__ void pack_offsets_kernel(const float* __restrict__ offset_r, const float* __restrict__ offset_t, float* __restrict__ packed, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // one (window, sensor) pair
    if (i >= n * kSensors) return;
    for (int e = 0; e < 9; ++e) packed[(int64_t)i * 12 + e] = offset_r[(int64_t)i * 9 + e];
    for (int e = 0; e < 3; ++e) packed[(int64_t)i * 12 + 9 + e] = offset_t[(int64_t)i * 3 + e];
}

// ---------------------------------------------------------------------------------------------
// CTA-cooperative mapping: a CTA owns kFramesPerCta frames whose scratch state lives in shared memory, and every
// phase of frame_math.h is flattened over (frame, item) across all threads of the CTA.  Narrow phases (the
// 3-row kinematic chain, the 12 sensors) then occupy one or two warps for ALL frames of the CTA instead of a
// few lanes in every warp, which is what the one-warp-per-frame mapping wasted most of its issue slots on.
// (kFramesPerCta, CTAs per SM) is a template parameter pair: more frames in flight per SM hide the latency of the narrow
// phases (the kinematic chains, the 12 sensor frames), fewer threads per frame make the wide phases longer.
constexpr int kMainThreads = 256;

// flattened (frame, item) loop over threads [t0, t0 + nt) of the CTA
#define EMPOSE_FOR_ITEMS_ON(t0, nt, n_items, f, i)                                                       \
    for (int _idx = (int)threadIdx.x - (t0), _n = (n_items), f = _idx / _n, i = _idx - f * _n;            \
         _idx >= 0 && _idx < nf * _n; _idx += (nt), f = _idx / _n, i = _idx - f * _n)
#define EMPOSE_FOR_ITEMS(n_items, f, i) EMPOSE_FOR_ITEMS_ON(0, kMainThreads, n_items, f, i)
// frame-grouped mapping for wide phases whose item count is a runtime value: kMainThreads / kFramesPerCta threads
// per frame, no integer division per item
#define EMPOSE_FOR_FRAME_ITEMS(n_items, f, i)                                                   \
    for (int f = threadIdx.x / kGroup, i = threadIdx.x % kGroup, _n = (n_items); f < nf && i < _n; i += kGroup)
static_assert(kMainThreads % 32 == 0, "whole warps");
// items of the serial kinematic chains, on the last warp of the CTA
#define EMPOSE_FOR_ITEMS_CHAIN(n_items, f, i) EMPOSE_FOR_ITEMS_ON(kMainThreads - 32, 32, n_items, f, i)

// optional phase timing (development aid): thread 0 of one mid-grid CTA stores clock64() after every barrier
#define EMPOSE_TICK(k) do { if (p.ticks && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) p.ticks[k] = clock64(); } while (0)

// Phases in the "one thread owns an item of EVERY frame of the CTA" form: the model constants (index ranges) are read
// once per CTA instead of once per frame.
// dE/dA_j = sum of its chunks (item_skin_bwd_reduce): thread (j, e), 22 x 12 items
template <int VP>
__device__ __forceinline__ void cta_skin_bwd_reduce(const SubModel& m, FrameState<float, VP>* st, int nf) {
    for (int it = threadIdx.x; it < kJoints * 12; it += kMainThreads) {
        const int j = it / 12, e = it - j * 12;
        const int c0 = __ldg(m.jvj_ptr + j), c1 = __ldg(m.jvj_ptr + j + 1);
        constexpr int kFrameFloats = (int)(sizeof(FrameState<float, VP>) / sizeof(float));
        const float* src = &st[0].dav[c0][e];
        float* dst = e < 9 ? &st[0].dar[j][e] : &st[0].dat[j][e - 9];
        for (int f = 0; f < nf; ++f, src += kFrameFloats, dst += kFrameFloats) {
            float acc = 0.0f;
            for (int c = 0; c < c1 - c0; ++c) acc += src[c * 12];
            *dst = acc;
        }
    }
}
// dE/dvp rows -> global memory, 16 bytes at a time (columns beyond 3 n_verts are zero: they are K padding of the
// transposed pose-blend GEMM)
template <int VP>
__device__ __forceinline__ void cta_store_dvp(const SubModel& m, FrameState<float, VP>* st, float* dvp, int64_t row0, int nf, int mode) {
    const int q = m.vp_dim / 4, nv3 = m.n_verts * 3;
    for (int idx = threadIdx.x; idx < nf * q; idx += kMainThreads) {
        const int f = idx / q, t = idx - f * q;
        float4 v = reinterpret_cast<const float4*>(st[f].dx)[t];
        const int i = t * 4;
        v.x = i < nv3 ? grad_operand_f32(v.x, mode) : 0.0f;
        v.y = i + 1 < nv3 ? grad_operand_f32(v.y, mode) : 0.0f;
        v.z = i + 2 < nv3 ? grad_operand_f32(v.z, mode) : 0.0f;
        v.w = i + 3 < nv3 ? grad_operand_f32(v.w, mode) : 0.0f;
        if (mode == OPERAND_F16) {
            const __half2 a = __floats2half2_rn(v.x * kDvpScale, v.y * kDvpScale), b = __floats2half2_rn(v.z * kDvpScale, v.w * kDvpScale);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b);
            reinterpret_cast<uint2*>(reinterpret_cast<__half*>(dvp) + (row0 + f) * m.vp_dim)[t] = pk;
        } else {
            reinterpret_cast<float4*>(dvp + (row0 + f) * m.vp_dim)[t] = v;
        }
    }
}

// The GENERAL sub-model kernel: any sensor neighbourhood (boundary or non-manifold rings, more than 8 skinning joints per
// ring), index-table driven, per-frame state in shared memory.  Sub-models in fan form -- every closed manifold mesh,
// i.e. SMPL-H -- run fan_kernel.cu instead; this one is the fallback and the A/B reference (EMPOSE_MAIN_GENERAL=1).
// Inputs and outputs are those of the fan kernel: blended rest vertices and rest joints in, dE/dvp and dE/dJ out.
template <int VP, int kFramesPerCta, int kMainCtasPerSm>
__global__ void __launch_bounds__(kMainThreads, kMainCtasPerSm) main_kernel(MainParams p) {
    constexpr int kGroup = kMainThreads / kFramesPerCta;     // threads per frame in the frame-grouped mapping
    static_assert(3 * kFramesPerCta <= 32, "the serial chains of all frames of a CTA run on one warp");
    extern __shared__ __align__(16) uint8_t smem_main[];
    FrameState<float, VP>* st = reinterpret_cast<FrameState<float, VP>*>(smem_main);
    const SubModel& m = p.sub;
    const int64_t row0 = (int64_t)blockIdx.x * kFramesPerCta;
    const int nf = (int)min((int64_t)kFramesPerCta, (int64_t)p.R - row0);
    const bool static_tree = p.static_tree != 0;
    EMPOSE_TICK(0);

    {   // the rows of the CTA's frames are contiguous in global memory: flat, coalesced copies
        const float* th = p.theta + row0 * kPoseDim;
        for (int idx = threadIdx.x; idx < nf * kPoseDim; idx += kMainThreads) { const int f = idx / kPoseDim; st[f].theta[idx - f * kPoseDim] = th[idx]; }
        const float* jr = p.jrest + row0 * kJrestLd;
        for (int idx = threadIdx.x; idx < nf * kJrestLd; idx += kMainThreads) {
            const int f = idx / kJrestLd, c = idx - f * kJrestLd;
            if (c < kPoseDim) (&st[f].jrest[0][0])[c] = jr[idx];
        }
        const int q = m.vp_dim / 4;
        for (int idx = threadIdx.x; idx < nf * q; idx += kMainThreads) {
            const int f = idx / q, t = idx - f * q;
            reinterpret_cast<float4*>(st[f].vp)[t] = __ldg(reinterpret_cast<const float4*>(p.vp + (row0 + f) * m.vp_dim) + t);
        }
    }
    __syncthreads();
    EMPOSE_TICK(1);
    EMPOSE_FOR_ITEMS(kJoints, f, i) item_rodrigues(st[f], i);
    __syncthreads();
    EMPOSE_TICK(2);
    if (threadIdx.x >= kMainThreads - 32) {
        if (static_tree) { EMPOSE_FOR_ITEMS_CHAIN(3, f, i) item_chain_static(st[f], i); }
        else { EMPOSE_FOR_ITEMS_CHAIN(3, f, i) item_chain(m, st[f], i); }
    }
    __syncthreads();
    EMPOSE_TICK(3);
    EMPOSE_FOR_FRAME_ITEMS(m.n_verts, f, i) item_skin(m, st[f], i);
    __syncthreads();
    EMPOSE_TICK(4);
    // sensor-major inputs -> the [12][3] / [12][9] views frame_math.h expects
    __shared__ float io[kFramesPerCta][2][144];            // [frame][offsets | measurement]: R (108) then t / position (36)
    EMPOSE_FOR_FRAME_ITEMS(144, f, i) {
        const int64_t row = row0 + f;
        const int64_t orow = row / p.rows_per_offset;
        const int sn = i / 12, e = i - sn * 12;
        if (p.offsets) io[f][0][e < 9 ? sn * 9 + e : 108 + sn * 3 + e - 9] = p.offsets[orow * 144 + i];
        else io[f][0][i] = i < 108 ? p.offset_r[orow * 108 + i] : p.offset_t[orow * 36 + i - 108];
        if (p.meas) io[f][1][e < 3 ? 108 + sn * 3 + e : sn * 9 + e - 3] = p.meas[row * 144 + i];
    }
    __syncthreads();
    if (m.max_degree <= kSplitDegree) {
        EMPOSE_FOR_FRAME_ITEMS(kSensors * m.max_degree, f, i) item_sensor_faces(m, st[f], i);
        __syncthreads();
    EMPOSE_TICK(5);
        EMPOSE_FOR_ITEMS(kSensors, f, i)
            item_sensor_frames(m, st[f], &io[f][0][0], &io[f][0][108], &io[f][1][108], &io[f][1][0], p.spec, p.want_grad != 0, i);
        __syncthreads();
    EMPOSE_TICK(6);
        if (p.want_grad) {
            EMPOSE_FOR_FRAME_ITEMS(kSensors * m.max_degree, f, i) item_sensor_face_grads(m, st[f], i);
            __syncthreads();
            EMPOSE_FOR_FRAME_ITEMS(m.n_verts, f, i) item_sensor_gather(m, st[f], i);
        }
    } else {
        EMPOSE_FOR_ITEMS(kSensors, f, i)
            item_sensors(m, st[f], &io[f][0][0], &io[f][0][108], &io[f][1][108], &io[f][1][0], p.spec, p.want_grad != 0, i);
    }
    __syncthreads();
    EMPOSE_TICK(7);
    if (p.sensor_pos) EMPOSE_FOR_ITEMS(36, f, i) p.sensor_pos[(row0 + f) * 36 + i] = st[f].sensor_pos[i / 3][i % 3];
    if (p.sensor_ori) EMPOSE_FOR_FRAME_ITEMS(108, f, i) p.sensor_ori[(row0 + f) * 108 + i] = st[f].sensor_ori[i / 9][i % 9];
    if (p.joints) EMPOSE_FOR_FRAME_ITEMS(kPoseDim, f, i) p.joints[(row0 + f) * kPoseDim + i] = st[f].gpos[i / 3][i % 3];
    if (!p.want_grad) return;
    const bool joint_up = p.joints_gt != nullptr;
    if (joint_up) EMPOSE_FOR_ITEMS(kJoints, f, i) item_joint_residual(st[f], p.joints_gt + (row0 + f) * kPoseDim, p.joint_weight, i);
    __syncthreads();
    EMPOSE_TICK(8);

    EMPOSE_FOR_FRAME_ITEMS(m.n_vj, f, i) item_skin_bwd_chunks(m, st[f], i);
    __syncthreads();
    EMPOSE_TICK(9);
    cta_skin_bwd_reduce(m, st, nf);
    EMPOSE_FOR_FRAME_ITEMS(m.n_verts, f, i) item_skin_bwd_verts(m, st[f], i);
    __syncthreads();
    EMPOSE_TICK(10);
    if (threadIdx.x >= kMainThreads - 32) {
        if (static_tree) { EMPOSE_FOR_ITEMS_CHAIN(3, f, i) item_chain_bwd_static(st[f], i, joint_up); }
        else { EMPOSE_FOR_ITEMS_CHAIN(3, f, i) item_chain_bwd(m, st[f], i, joint_up); }
    }
    cta_store_dvp(m, st, p.dvp, row0, nf, p.round_out);
    __syncthreads();
    EMPOSE_TICK(11);
    EMPOSE_FOR_FRAME_ITEMS(kJoints * 12, f, i) item_chain_bwd_local(m, st[f], i, joint_up);
    __syncthreads();
    EMPOSE_TICK(12);
    EMPOSE_FOR_ITEMS(kJoints, f, i)
        item_finish_theta(st[f], p.coef[row0 + f], (const float*)nullptr, p.gtheta_part + (row0 + f) * kPoseDim, i);
    EMPOSE_FOR_FRAME_ITEMS(p.dj_ld, f, i) {
        const float v = i < kPoseDim ? st[f].dj[i / 3][i % 3] : 0.0f;
        if (p.round_out == OPERAND_F16) reinterpret_cast<__half*>(p.dj)[(row0 + f) * p.dj_ld + i] = __float2half_rn(v * kDvpScale);
        else p.dj[(row0 + f) * p.dj_ld + i] = grad_operand_f32(v, p.round_out);
    }
    EMPOSE_TICK(13);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) post_kernel(PostParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)p.R * 32) return;
    const int64_t row = i >> 5;
    const int j = (int)(i & 31);
    const int64_t xg = row * p.iter_stride + p.in_size + 76;      // first gradient column of this row in xiter
    if (j < kJoints) {
        float g[3] = {p.gtheta_part[row * kPoseDim + j * 3], p.gtheta_part[row * kPoseDim + j * 3 + 1],
                      p.gtheta_part[row * kPoseDim + j * 3 + 2]};
        if (j > 0) {
            const float c = p.coef[row];
            float r[3] = {p.theta[row * kPoseDim + j * 3], p.theta[row * kPoseDim + j * 3 + 1], p.theta[row * kPoseDim + j * 3 + 2]};
            float dR[9];
            const float* src = p.dpf + row * kPoseFeatPad + (j - 1) * 9;
#pragma unroll
            for (int e = 0; e < 9; ++e) dR[e] = src[e] * c;
            rodrigues_bwd(r, dR, g);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            if (p.xiter) store_operand(p.xiter, xg + j * 3 + e, g[e], p.operand_mode);
            if (p.g_theta_out) p.g_theta_out[row * kPoseDim + j * 3 + e] = g[e];
        }
    } else if (j < kJoints + kBetas) {
        const int k = j - kJoints;
        const float g = p.coef[row] * p.dpf[row * kPoseFeatPad + kFeatBeta + k];
        if (p.xiter) store_operand(p.xiter, xg + kPoseDim + k, g, p.operand_mode);
        if (p.g_beta_out) p.g_beta_out[row * kBetas + k] = g;
    }
}

__global__ void gather_last_kernel(const float* __restrict__ seq, float* __restrict__ out, int B, int F, int H, int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H) return;
    const int64_t b = i / H;
    const int u = (int)(i % H);
    out[i] = load_operand(seq, (b * F + (F - 1)) * H + u, mode);
}

__global__ void to_operand_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_operand(dst, i, src[i], mode);
}

__global__ void convert_2d_kernel(const float* __restrict__ src, int64_t src_ld, int64_t rows, int cols, float* __restrict__ dst,
                                  int64_t dst_ld, int mode, int to_operand) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    if (to_operand) store_operand(dst, r * dst_ld + c, src[r * src_ld + c], mode);
    else dst[r * dst_ld + c] = load_operand(src, r * src_ld + c, mode);
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

int launch_prepare(const PrepareParams& p, cudaStream_t s) {
    if (p.operand_mode == OPERAND_F16 && p.in_size % 8 == 0 && p.in_size <= 144 && p.in_stride % 8 == 0 && p.iter_stride % 8 == 0) {
        prepare_rows_kernel<<<blocks_for(p.R, kPrepRows), 256, 0, s>>>(p);
        EMPOSE_CUDA_TRY(cudaGetLastError());
        return EMPOSE_OK;
    }
    prepare_kernel<<<blocks_for((int64_t)p.R * 144, 256), 256, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_update(const UpdateParams& p, cudaStream_t s) {
    // the staging tile holds kUpdateGroup feature rows: half the bytes with fp16 operands, so eight CTAs fit an SM instead of six
    const size_t smem = (size_t)kUpdateGroup * (p.pf_split == OPERAND_F16 ? p.pf_stride / 2 : p.pf_stride) * sizeof(float);
    static size_t configured = 0;
    if (smem > 40 * 1024 && smem > configured) {
        EMPOSE_CUDA_TRY(cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    update_kernel<<<p.B, kUpdateThreads, smem, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_pose_features(const float* theta, const float* beta, float* pf, int pf_stride, int pf_split, int R, cudaStream_t s) {
    pose_feature_kernel<<<blocks_for((int64_t)R * (kJoints - 1), 256), 256, 0, s>>>(theta, beta, pf, pf_stride, pf_split, R);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_pack_offsets(const float* offset_r, const float* offset_t, float* packed, int n, cudaStream_t s) {
    pack_offsets_kernel<<<blocks_for((int64_t)n * kSensors, 128), 128, 0, s>>>(offset_r, offset_t, packed, n);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

template <int VP, int FPC, int CPS>
int launch_main_variant(const MainParams& p, cudaStream_t s) {
    static bool configured = false;
    const size_t smem = sizeof(FrameState<float, VP>) * FPC;      // (+ FPC * 1152 B static)
    if (!configured) {
        EMPOSE_CUDA_TRY(cudaFuncSetAttribute(main_kernel<VP, FPC, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    main_kernel<VP, FPC, CPS><<<blocks_for(p.R, FPC), kMainThreads, smem, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

DebugOptions& debug_options() {
    static DebugOptions opt = [] {
        DebugOptions o;
        const char* e = getenv("EMPOSE_MAIN_GENERAL");
        o.main_general = e ? atoi(e) : 0;
        e = getenv("EMPOSE_FAN_VARIANT");
        o.fan_variant = e ? atoi(e) : 0;
        e = getenv("EMPOSE_LSTM_PERSISTENT");
        o.lstm_persistent = e ? atoi(e) : 1;
        e = getenv("EMPOSE_BLEND_FP16");
        o.blend_fp16 = e ? atoi(e) : 1;
        return o;
    }();
    return opt;
}

int launch_main(const MainParams& p, cudaStream_t s) {
    if (p.fan.ok && !debug_options().main_general) return launch_main_fan(p, s);
    if (p.sub.vp_dim <= 288) return launch_main_variant<288, 5, 3>(p, s);
    if (p.sub.vp_dim <= kMaxVp) return launch_main_variant<kMaxVp, 3, 3>(p, s);
    set_last_error("sensor sub-mesh too large (more than 144 vertices)");
    return EMPOSE_E_ARG;
}

int launch_post(const PostParams& p, cudaStream_t s) {
    post_kernel<<<blocks_for((int64_t)p.R * 32, 256), 256, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_gather_last(const float* seq, float* out, int B, int F, int H, int operand_mode, cudaStream_t s) {
    gather_last_kernel<<<blocks_for((int64_t)B * H, 256), 256, 0, s>>>(seq, out, B, F, H, operand_mode);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_to_operand_2d(const float* src, int64_t src_ld, int64_t rows, int cols, float* dst, int64_t dst_ld, int operand_mode,
                         cudaStream_t s) {
    convert_2d_kernel<<<blocks_for(rows * cols, 256), 256, 0, s>>>(src, src_ld, rows, cols, dst, dst_ld, operand_mode, 1);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_from_operand_2d(const float* src, int64_t src_ld, int64_t rows, int cols, float* dst, int64_t dst_ld, int operand_mode,
                           cudaStream_t s) {
    convert_2d_kernel<<<blocks_for(rows * cols, 256), 256, 0, s>>>(src, src_ld, rows, cols, dst, dst_ld, operand_mode, 0);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_to_operand(const float* src, float* dst, int64_t n, int operand_mode, cudaStream_t s) {
    to_operand_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, dst, n, operand_mode);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

}  // namespace empose
