// The per-frame SMPL-H sub-model pass of the LGD loop in fan form (fan_math.h): forward + hand-derived reverse pass.
//
// Replaces, per frame, the reference's full-mesh BodyModel call (empose/bodymodels/smpl.py:121), the sensor frames
// (empose/data/virtual_sensors.py:85-96), the offsets (empose/nn/models.py:478-479), reconstruction_loss
// (empose/nn/loss.py:23-41) and autograd's backward through all of it (models.py:576-579).
//
// Mapping.  A CTA owns G consecutive frames.  The wide part of the work -- skinning a sensor's ring, its frame, the
// residual, and the reverse pass down to dE/dvp and the dE/dA partial sums -- is one lane per (frame, sensor) with
// everything in registers (12 G lanes, no shared-memory state, no index tables: fan_sensor_item).  The narrow,
// serial parts run on the per-frame JointState in shared memory: Rodrigues (22 G items), the kinematic chain (3 G
// lanes of the last warp, one row of every transform each), the fixed-order reduction of the partial sums (no
// atomics: results are bit-reproducible whatever batch a window is part of), the reverse sweep (3 G lanes), the local
// gradients (264 G items) and Rodrigues' reverse (22 G items).  Seven CTA barriers per launch; ~6 KB of shared
// memory per frame, so four CTAs of eight frames share an SM and hide each other's narrow phases.
//
// The shape blend, the template, the pose blend and the rest joints all come from ONE tensor-core contraction
// before this kernel (`vp`, `jrest`), and dE/dvp, dE/dJ go back through its transpose after it (model.cu).
#include <stdlib.h>

#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_kernels.h"

namespace empose {
namespace {

template <int SLOTS, int MAXD, int G>
struct FanCfg {
    static constexpr int kRing = MAXD + 1;
    static constexpr int kLanes = kSensors * G;                       // (frame, sensor) items
    static constexpr int kThreads = ((kLanes + 31) / 32) * 32;
    static constexpr int kVpVec = (kRing * 3 + 3) / 4;                // float4 loads of one ring
    static constexpr int kBlockVec = SLOTS * 3 / 4;                   // float4 per sensor block of vp / dvp
    static constexpr int kChainThreads = ((3 * G + 31) / 32) * 32;    // the serial chains run on the last warp(s)
    static_assert(kChainThreads <= kThreads, "chain lanes must fit the CTA");
};

// optional phase timing (EMPOSE_MAIN_TICKS): thread 0 of one mid-grid CTA stores clock64() at the phase boundaries
#define FAN_TICK(k) do { if (p.ticks && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) p.ticks[k] = clock64(); } while (0)

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ float grad_operand_f32(float x, int mode) { return mode == OPERAND_TF32 ? round_tf32(x) : x; }

template <int SLOTS, int MAXD, int G, int CTAS>
__global__ void __launch_bounds__((FanCfg<SLOTS, MAXD, G>::kThreads), CTAS) fan_kernel(MainParams p, int frame_bytes) {
    using Cfg = FanCfg<SLOTS, MAXD, G>;
    constexpr int NT = Cfg::kThreads;
    constexpr int RING = Cfg::kRing;
    extern __shared__ __align__(16) uint8_t smem_fan[];
    auto state = [&](int f) -> JointState<float>& { return *reinterpret_cast<JointState<float>*>(smem_fan + (size_t)f * frame_bytes); };
    auto var_of = [&](int f) -> float* { return reinterpret_cast<float*>(smem_fan + (size_t)f * frame_bytes + sizeof(JointState<float>)); };

    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * G;
    const int nf = (int)min((int64_t)G, (int64_t)p.R - row0);
    const int fs = tid / kSensors, s = tid - fs * kSensors;          // this lane's (frame, sensor) item
    const bool grad = p.want_grad != 0;
    // nobody asked for the sensors (the final evaluation of a pass without marker histories): joints only -- the (frame, sensor)
    // items, and with them the ring / offset / measurement loads, are skipped
    const bool item = tid < kSensors * nf && (grad || p.sensor_pos != nullptr || p.sensor_ori != nullptr);

    // ---- P0: inputs.  The lane's ring, offsets and measurement go straight to registers; their latency hides behind
    // ---- the joint phases.
    FAN_TICK(0);
    float vp[RING * 3], off[12], meas[12];
    if (item) {
        const int64_t row = row0 + fs;
        const float4* src = reinterpret_cast<const float4*>(p.vp + row * p.sub.vp_dim + s * (SLOTS * 3));
        float4 q[Cfg::kVpVec];
#pragma unroll
        for (int i = 0; i < Cfg::kVpVec; ++i) q[i] = __ldg(src + i);
#pragma unroll
        for (int i = 0; i < RING * 3; ++i) {
            const float4 v = q[i / 4];
            vp[i] = (i % 4 == 0) ? v.x : (i % 4 == 1) ? v.y : (i % 4 == 2) ? v.z : v.w;
        }
        const int64_t orow = row / p.rows_per_offset;
        if (p.offsets) {
            const float4* o4 = reinterpret_cast<const float4*>(p.offsets + (orow * kSensors + s) * 12);
            const float4 a = __ldg(o4), b = __ldg(o4 + 1), c = __ldg(o4 + 2);
            off[0] = a.x; off[1] = a.y; off[2] = a.z; off[3] = a.w; off[4] = b.x; off[5] = b.y; off[6] = b.z; off[7] = b.w;
            off[8] = c.x; off[9] = c.y; off[10] = c.z; off[11] = c.w;
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) off[i] = __ldg(p.offset_r + orow * 108 + s * 9 + i);
#pragma unroll
            for (int i = 0; i < 3; ++i) off[9 + i] = __ldg(p.offset_t + orow * 36 + s * 3 + i);
        }
        if (grad) {
            const float4* m4 = reinterpret_cast<const float4*>(p.meas + (row * kSensors + s) * 12);
            const float4 a = __ldg(m4), b = __ldg(m4 + 1), c = __ldg(m4 + 2);
            meas[0] = a.x; meas[1] = a.y; meas[2] = a.z; meas[3] = a.w; meas[4] = b.x; meas[5] = b.y; meas[6] = b.z; meas[7] = b.w;
            meas[8] = c.x; meas[9] = c.y; meas[10] = c.z; meas[11] = c.w;
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) meas[i] = 0.0f;
        }
    }
    {   // rows of the CTA's frames are contiguous in global memory: flat, coalesced copies, straight into shared memory
        // (cp.async: all of a thread's ~11 words are in flight at once.  Through registers the compiler kept each load next to
        //  its store, a full memory round trip per loop iteration: the phase took ~8000 of the CTA's ~46000 cycles.)
        const float* th = p.theta + row0 * kPoseDim;
        for (int idx = tid; idx < nf * kPoseDim; idx += NT) { const int f = idx / kPoseDim; cp_async_4(&state(f).theta[idx - f * kPoseDim], th + idx); }
        const float* jr = p.jrest + row0 * kJrestLd;
        for (int idx = tid; idx < nf * kJrestLd; idx += NT) {
            const int f = idx / kJrestLd, c = idx - f * kJrestLd;
            if (c < kPoseDim) cp_async_4(&state(f).jrest[0][0] + c, jr + idx);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    FAN_TICK(1);
    // ---- P1: joint rotations ----
    for (int idx = tid; idx < nf * kJoints; idx += NT) { const int f = idx / kJoints; jt_rodrigues(state(f), idx - f * kJoints); }
    __syncthreads();
    FAN_TICK(2);
    // ---- P2: kinematic chain, three lanes per frame on the last warp ----
    if (tid >= NT - Cfg::kChainThreads) {
        const int l = tid - (NT - Cfg::kChainThreads);
        if (l < 3 * nf) {
            if (p.static_tree) jt_chain_static(state(l / 3), l % 3);
            else jt_chain(p.sub.parents, state(l / 3), l % 3);
        }
    }
    __syncthreads();
    FAN_TICK(3);
    // ---- P3: the (frame, sensor) items ----
    if (item) {
        const int64_t row = row0 + fs;
        float out_pos[3], out_ori[9], dvp[RING * 3];
        fan_sensor_item<float, SLOTS, MAXD>(p.fan, s, &state(fs).A[0][0], vp, off, meas, p.spec, grad, out_pos, out_ori, dvp, var_of(fs));
        if (p.sensor_pos) {
#pragma unroll
            for (int i = 0; i < 3; ++i) p.sensor_pos[row * 36 + s * 3 + i] = out_pos[i];
        }
        if (p.sensor_ori) {
#pragma unroll
            for (int i = 0; i < 9; ++i) p.sensor_ori[row * 108 + s * 9 + i] = out_ori[i];
        }
        if (grad) {         // the whole sensor block, padding columns as zeros (K of the transposed blend GEMM)
            const int mode = p.round_out;
            auto val = [&](int i) { return i < RING * 3 ? grad_operand_f32(dvp[i < RING * 3 ? i : 0], mode) : 0.0f; };
            if (mode == OPERAND_F16) {      // fp16 elements of value * kDvpScale, four per store
                uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.dvp) + row * p.sub.vp_dim + s * (SLOTS * 3));
#pragma unroll
                for (int i = 0; i < Cfg::kBlockVec; ++i) {
                    const __half2 a = __floats2half2_rn(val(4 * i) * kDvpScale, val(4 * i + 1) * kDvpScale);
                    const __half2 b = __floats2half2_rn(val(4 * i + 2) * kDvpScale, val(4 * i + 3) * kDvpScale);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b);
                    dst[i] = pk;
                }
            } else {
                float4* dst = reinterpret_cast<float4*>(p.dvp + row * p.sub.vp_dim + s * (SLOTS * 3));
#pragma unroll
                for (int i = 0; i < Cfg::kBlockVec; ++i) dst[i] = make_float4(val(4 * i), val(4 * i + 1), val(4 * i + 2), val(4 * i + 3));
            }
        }
    }
    if (p.joints) {
        float* dst = p.joints + row0 * kPoseDim;
        for (int idx = tid; idx < nf * kPoseDim; idx += NT) { const int f = idx / kPoseDim; dst[idx] = (&state(f).gpos[0][0])[idx - f * kPoseDim]; }
    }
    if (!grad) return;
    FAN_TICK(4);
    __syncthreads();
    FAN_TICK(5);
    // ---- P4: dE/dA_j = sum of the partials naming joint j, fixed order; a thread owns (j, e) of EVERY frame, so the
    // ---- lists are read once per CTA.  Training: the FK-loss upstream replaces the posed joints in place.
    const bool joint_up = p.joints_gt != nullptr;
    jt_reduce_frames<float>(p.fan, state, var_of, nf, tid, NT);
    if (joint_up)
        for (int idx = tid; idx < nf * kJoints; idx += NT) {
            const int f = idx / kJoints;
            jt_joint_residual(state(f), p.joints_gt + (row0 + f) * kPoseDim, p.joint_weight, idx - f * kJoints);
        }
    __syncthreads();
    FAN_TICK(6);
    // ---- P5: reverse sweep of the chain ----
    if (tid >= NT - Cfg::kChainThreads) {
        const int l = tid - (NT - Cfg::kChainThreads);
        if (l < 3 * nf) {
            if (p.static_tree) jt_chain_bwd_static(state(l / 3), l % 3, joint_up);
            else jt_chain_bwd(p.sub.parents, state(l / 3), l % 3, joint_up);
        }
    }
    __syncthreads();
    FAN_TICK(7);
    // ---- P6: local gradients dE/dR_j, dE/dJ_j (over the dead partial sums) ----
    jt_local_frames<float>(p.sub.parents, state, var_of, nf, tid, NT, joint_up);
    __syncthreads();
    FAN_TICK(8);
    // ---- P7: outputs ----
    for (int idx = tid; idx < nf * kJoints; idx += NT) {
        const int f = idx / kJoints;
        jt_finish_theta(state(f), var_of(f), p.coef[row0 + f], p.gtheta_part + (row0 + f) * kPoseDim, idx - f * kJoints);
    }
    {
        const int ld = p.dj_ld;
        if (p.round_out == OPERAND_F16 && (ld & 3) == 0) {      // four fp16 values per store (rows of the CTA's frames are contiguous)
            uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.dj) + row0 * ld);
            for (int q = tid; q < nf * ld / 4; q += NT) {
                const int f = (4 * q) / ld, c = 4 * q - f * ld;
                const float* v = var_of(f) + kJoints * 9;
                float x[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = c + k < kPoseDim ? v[c + k] * kDvpScale : 0.0f;
                const __half2 a = __floats2half2_rn(x[0], x[1]), b = __floats2half2_rn(x[2], x[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b);
                dst[q] = pk;
            }
        } else
        for (int idx = tid; idx < nf * ld; idx += NT) {
            const int f = idx / ld, c = idx - f * ld;
            const float v = c < kPoseDim ? var_of(f)[kJoints * 9 + c] : 0.0f;
            if (p.round_out == OPERAND_F16) reinterpret_cast<__half*>(p.dj)[row0 * ld + idx] = __float2half_rn(v * kDvpScale);
            else p.dj[row0 * ld + idx] = grad_operand_f32(v, p.round_out);
        }
    }
    FAN_TICK(9);
}

template <int SLOTS, int MAXD, int G, int CTAS>
int launch_variant(const MainParams& p, cudaStream_t s) {
    using Cfg = FanCfg<SLOTS, MAXD, G>;
    static int configured_bytes = -1;
    const int frame_bytes = (int)sizeof(JointState<float>) + ((fan_var_floats(p.fan.n_part) + 3) / 4) * 16;
    const int smem = frame_bytes * G;
    if (smem > 227 * 1024) { set_last_error("fan kernel: shared memory per CTA out of range"); return EMPOSE_E_ARG; }
    if (configured_bytes < smem) {
        EMPOSE_CUDA_TRY(cudaFuncSetAttribute(fan_kernel<SLOTS, MAXD, G, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_bytes = smem;
    }
    const unsigned grid = (unsigned)((p.R + G - 1) / G);
    fan_kernel<SLOTS, MAXD, G, CTAS><<<grid, Cfg::kThreads, smem, s>>>(p, frame_bytes);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

}  // namespace

int launch_main_fan(const MainParams& p, cudaStream_t s) {
    // option fan_variant selects (frames per CTA, CTAs per SM) for the common 8-slot layout (experiments)
    const int variant = debug_options().fan_variant;
    const FanModel& fm = p.fan;
    if (!fm.ok || (fm.slots != 8 && fm.slots != 12) || fm.max_deg >= fm.slots || fm.n_part > kMaxPartials || p.sub.vp_dim < kSensors * fm.slots * 3) {
        set_last_error("fan kernel: sub-model is not in fan form");
        return EMPOSE_E_ARG;
    }
    if (fm.slots == 8 && fm.max_deg <= 6) {
        switch (variant) {
            case 1: return launch_variant<8, 6, 16, 2>(p, s);
            case 2: return launch_variant<8, 6, 8, 3>(p, s);
            case 3: return launch_variant<8, 6, 5, 6>(p, s);
            case 4: return launch_variant<8, 6, 7, 5>(p, s);
            case 5: return launch_variant<8, 6, 10, 3>(p, s);
            case 6: return launch_variant<8, 6, 5, 5>(p, s);
            default: return launch_variant<8, 6, 8, 4>(p, s);
        }
    }
    if (fm.slots == 8) return launch_variant<8, 7, 8, 4>(p, s);
    return launch_variant<12, 11, 8, 2>(p, s);
}

}  // namespace empose
