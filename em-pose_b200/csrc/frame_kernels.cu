// Per-frame CUDA kernels of the LGD loop: input assembly, estimate update + pose features, the SMPL
// sub-model forward / reverse pass (one warp per frame, per-frame state in shared memory, arithmetic
// in frame_math.h) and the gradient-feature finish.  Reference call sites are cited in
// frame_kernels.h and frame_math.h.
#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_kernels.h"

namespace empose {

namespace {

__device__ __forceinline__ float maybe_round(float x, int round_out) { return round_out ? round_tf32(x) : x; }

// pose features of one joint: vec(R - I), optionally split into a tf32 value and its tf32 residual
__device__ __forceinline__ void write_pose_features(const float* r, float* dst, int split) {
    float R[9];
    rodrigues_fwd(r, R);
#pragma unroll
    for (int e = 0; e < 9; ++e) {
        const float v = R[e] - ((e % 4 == 0) ? 1.0f : 0.0f);
        if (split) {
            const float hi = round_tf32(v);
            dst[e] = hi;
            dst[kPoseFeatPad + e] = round_tf32(v - hi);
        } else {
            dst[e] = v;
        }
    }
}

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
    } else {
        const int e = c - 36;
        v = p.marker_oris[(int64_t)row * 108 + e];
        const int slot = p.slot_of_sensor[e / 9];
        if (p.use_ori && slot >= 0) dst = p.n_pos + slot * 9 + e % 9;
    }
    p.meas[idx] = v;
    if (dst >= 0) {
        const float r = maybe_round(v, p.round_out);
        if (p.xin) p.xin[(int64_t)row * p.in_size + dst] = r;
        if (p.xiter) p.xiter[(int64_t)row * p.iter_in + dst] = r;
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

// ---------------------------------------------------------------------------------------------
// one CTA per window
__global__ void __launch_bounds__(128) update_kernel(UpdateParams p) {
    __shared__ float mean_db[kBetas];
    const int b = blockIdx.x;
    const int64_t row0 = (int64_t)b * p.F;
    if (p.average_shape) {
        if (threadIdx.x < kBetas) {
            float acc = 0.0f;
            for (int f = 0; f < p.F; ++f) acc += p.dbeta[(row0 + f) * kBetas + threadIdx.x];
            mean_db[threadIdx.x] = acc / (float)p.F;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < p.F * 76; i += blockDim.x) {
        const int f = i / 76, c = i % 76;
        const int64_t row = row0 + f;
        float v;
        if (c < kPoseDim) {
            const float d = p.dtheta[row * kPoseDim + c];
            v = p.first ? d : p.theta[row * kPoseDim + c] + p.step * d;
            p.theta[row * kPoseDim + c] = v;
            if (p.hist_pose) p.hist_pose[row * kPoseDim + c] = v;
        } else {
            const int k = c - kPoseDim;
            const float d = p.average_shape ? mean_db[k] : p.dbeta[row * kBetas + k];
            v = p.first ? d : p.beta[row * kBetas + k] + p.step * d;
            p.beta[row * kBetas + k] = v;
            if (p.hist_shape) p.hist_shape[row * kBetas + k] = v;
        }
        if (p.xiter) p.xiter[row * p.iter_in + p.in_size + c] = maybe_round(v, p.round_out);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.F * (kJoints - 1); i += blockDim.x) {
        const int f = i / (kJoints - 1), j = 1 + i % (kJoints - 1);
        const int64_t row = row0 + f;
        float r[3] = {p.theta[row * kPoseDim + j * 3], p.theta[row * kPoseDim + j * 3 + 1], p.theta[row * kPoseDim + j * 3 + 2]};
        write_pose_features(r, p.pf + row * p.pf_stride + (j - 1) * 9, p.pf_split);
    }
}

__global__ void __launch_bounds__(256) pose_feature_kernel(const float* __restrict__ theta, float* __restrict__ pf,
                                                           int pf_stride, int pf_split, int R) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)R * (kJoints - 1)) return;
    const int64_t row = i / (kJoints - 1);
    const int j = 1 + (int)(i % (kJoints - 1));
    float r[3] = {theta[row * kPoseDim + j * 3], theta[row * kPoseDim + j * 3 + 1], theta[row * kPoseDim + j * 3 + 2]};
    write_pose_features(r, pf + row * pf_stride + (j - 1) * 9, pf_split);
}

// ---------------------------------------------------------------------------------------------
constexpr int kFramesPerCta = 4;   // one warp per frame

template <int VP>
__global__ void __launch_bounds__(kFramesPerCta * 32) main_kernel(MainParams p) {
    extern __shared__ __align__(16) uint8_t smem_main[];
    FrameState<float, VP>* states = reinterpret_cast<FrameState<float, VP>*>(smem_main);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kFramesPerCta + warp;
    if (row >= p.R) return;                     // warps are independent: only __syncwarp below
    FrameState<float, VP>& st = states[warp];
    const SubModel& m = p.sub;

    for (int i = lane; i < kPoseDim; i += 32) st.theta[i] = p.theta[row * kPoseDim + i];
    if (lane < kBetas) st.beta[lane] = p.beta[row * kBetas + lane];
    __syncwarp();
    phase_setup(m, st, p.vp_off + row * m.vp_dim, lane, 32);
    __syncwarp();
    phase_chain(m, st, lane, 32);
    __syncwarp();
    phase_skin(m, st, lane, 32);
    __syncwarp();
    const int64_t orow = row / p.rows_per_offset;
    const float* meas = p.meas ? p.meas + row * 144 : nullptr;
    phase_sensors(m, st, p.offset_r + orow * 108, p.offset_t + orow * 36, meas, meas ? meas + 36 : nullptr, p.spec,
                  p.want_grad != 0, lane, 32);
    __syncwarp();
    if (p.sensor_pos)
        for (int i = lane; i < 36; i += 32) p.sensor_pos[row * 36 + i] = st.sensor_pos[i / 3][i % 3];
    if (p.sensor_ori)
        for (int i = lane; i < 108; i += 32) p.sensor_ori[row * 108 + i] = st.sensor_ori[i / 9][i % 9];
    if (p.joints)
        for (int i = lane; i < kPoseDim; i += 32) p.joints[row * kPoseDim + i] = st.gpos[i / 3][i % 3];
    if (!p.want_grad) return;

    phase_skin_bwd_joints(m, st, lane, 32);
    __syncwarp();
    phase_skin_bwd_verts(m, st, lane, 32);
    __syncwarp();
    {
        const int nv3 = m.n_verts * 3;
        float* dst = p.dvp + row * m.vp_dim;
        for (int i = lane; i < m.vp_dim; i += 32) dst[i] = i < nv3 ? maybe_round(st.dx[i], p.round_out) : 0.0f;
    }
    phase_shape_bwd_partial(m, st, lane, 32);
    phase_chain_bwd(m, st, lane, 32);
    __syncwarp();
    phase_chain_bwd_local(m, st, lane, 32);
    __syncwarp();
    phase_finish(m, st, p.coef[row], (const float*)nullptr, p.gtheta_part + row * kPoseDim, p.gbeta + row * kBetas, lane,
                 32);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) post_kernel(PostParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)p.R * 32) return;
    const int64_t row = i >> 5;
    const int j = (int)(i & 31);
    float* xg = p.xiter ? p.xiter + row * p.iter_in + p.in_size + 76 : nullptr;
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
            if (xg) xg[j * 3 + e] = maybe_round(g[e], p.round_out);
            if (p.g_theta_out) p.g_theta_out[row * kPoseDim + j * 3 + e] = g[e];
        }
    } else if (j < kJoints + kBetas) {
        const int k = j - kJoints;
        const float g = p.gbeta[row * kBetas + k];
        if (xg) xg[kPoseDim + k] = maybe_round(g, p.round_out);
        if (p.g_beta_out) p.g_beta_out[row * kBetas + k] = g;
    }
}

__global__ void gather_last_kernel(const float* __restrict__ seq, float* __restrict__ out, int B, int F, int H) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H) return;
    const int64_t b = i / H;
    const int u = (int)(i % H);
    out[i] = seq[(b * F + (F - 1)) * H + u];
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

int launch_prepare(const PrepareParams& p, cudaStream_t s) {
    prepare_kernel<<<blocks_for((int64_t)p.R * 144, 256), 256, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_update(const UpdateParams& p, cudaStream_t s) {
    update_kernel<<<p.B, 128, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_pose_features(const float* theta, float* pf, int pf_stride, int pf_split, int R, cudaStream_t s) {
    pose_feature_kernel<<<blocks_for((int64_t)R * (kJoints - 1), 256), 256, 0, s>>>(theta, pf, pf_stride, pf_split, R);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_main(const MainParams& p, cudaStream_t s) {
    const unsigned grid = blocks_for(p.R, kFramesPerCta);
    if (p.sub.vp_dim <= 256) {
        const size_t smem = sizeof(FrameState<float, 256>) * kFramesPerCta;
        main_kernel<256><<<grid, kFramesPerCta * 32, smem, s>>>(p);
    } else if (p.sub.vp_dim <= kMaxVp) {
        const size_t smem = sizeof(FrameState<float, kMaxVp>) * kFramesPerCta;
        static bool configured = false;
        if (!configured) {
            EMPOSE_CUDA_TRY(cudaFuncSetAttribute(main_kernel<kMaxVp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = true;
        }
        main_kernel<kMaxVp><<<grid, kFramesPerCta * 32, smem, s>>>(p);
    } else {
        set_last_error("sensor sub-mesh too large (more than 128 vertices)");
        return EMPOSE_E_ARG;
    }
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_post(const PostParams& p, cudaStream_t s) {
    post_kernel<<<blocks_for((int64_t)p.R * 32, 256), 256, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int launch_gather_last(const float* seq, float* out, int B, int F, int H, cudaStream_t s) {
    gather_last_kernel<<<blocks_for((int64_t)B * H, 256), 256, 0, s>>>(seq, out, B, F, H);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

}  // namespace empose
