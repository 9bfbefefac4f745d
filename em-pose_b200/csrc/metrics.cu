// Evaluation metrics on the device: what the reference's MetricsEngine.compute does per frame
// (empose/eval/metrics.py:183-241) -- two SMPL forward-kinematics passes (ground truth and prediction), per-joint
// Euclidean distances, the same after a per-frame Procrustes alignment with optimal scale (_procrustes, :19-66, a numpy
// SVD per frame on the host in the reference) and the geodesic angle between the global joint orientations obtained with
// a zero root (:229-238) -- as ONE kernel, one thread per frame.  Only the 22 body joints are needed, so the FK uses the
// folded joint regressor of the sub-model (J(beta) = j0 + jdirs beta) and never touches a vertex.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_math.h"
#include "metrics_math.h"
#include "model_internal.h"

namespace empose {
namespace {

__global__ void __launch_bounds__(128) metrics_kernel(MetricsParams p) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= p.R) return;
    float jg[kJoints][3], jh[kJoints][3];
    if (p.joints) {
        for (int i = 0; i < kPoseDim; ++i) { jg[i / 3][i % 3] = p.joints[(int64_t)row * kPoseDim + i]; jh[i / 3][i % 3] = p.joints_hat[(int64_t)row * kPoseDim + i]; }
    } else {
        float og[kJoints][9], oh[kJoints][9];
        fk_frame(p, p.pose + (int64_t)row * kPoseDim, p.shape + (int64_t)row * kBetas, jg, og);
        fk_frame(p, p.pose_hat + (int64_t)row * kPoseDim, p.shape_hat + (int64_t)row * kBetas, jh, oh);
        if (p.angle)
            for (int j = 1; j < kJoints; ++j) {      // geodesic angle of og^T oh (quaternion.rotation_intrinsic_distance), degrees
                if (p.angle_local) {                  // the joint's own rotations instead of the accumulated ones
                    rodrigues_fwd(p.pose + (int64_t)row * kPoseDim + j * 3, og[j]);
                    rodrigues_fwd(p.pose_hat + (int64_t)row * kPoseDim + j * 3, oh[j]);
                }
                float d[9];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) d[r * 3 + c] = og[j][r] * oh[j][c] + og[j][3 + r] * oh[j][3 + c] + og[j][6 + r] * oh[j][6 + c];
                const float cs = 0.5f * (d[0] + d[4] + d[8] - 1.0f);
                const float s0 = 0.5f * (d[7] - d[5]), s1 = 0.5f * (d[2] - d[6]), s2 = 0.5f * (d[3] - d[1]);
                p.angle[(int64_t)row * (kJoints - 1) + j - 1] = atan2f(sqrtf(s0 * s0 + s1 * s1 + s2 * s2), cs) * 57.29577951308232f;
            }
    }
    joint_distances(jg, jh, p.eucl + (int64_t)row * kJoints, p.eucl_pa + (int64_t)row * kJoints);
}

int launch_metrics(const MetricsParams& p, cudaStream_t s) {
    if (p.R <= 0) return EMPOSE_OK;
    metrics_kernel<<<(p.R + 127) / 128, 128, 0, s>>>(p);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

}  // namespace
}  // namespace empose

using namespace empose;

extern "C" {
#pragma GCC visibility push(default)

int empose_metrics_compute(empose_ief* ctx, const float* pose, const float* shape, const float* pose_hat, const float* shape_hat,
                           int32_t R, int32_t angle_local, float* eucl, float* eucl_pa, float* angle_deg, void* stream) {
    if (!ctx || !pose || !shape || !pose_hat || !shape_hat || !eucl || !eucl_pa || R < 0) { set_last_error("bad argument"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    MetricsParams p;
    memset(&p, 0, sizeof(p));
    p.j0 = ctx->sub.j0; p.jdirs = ctx->sub.jdirs; p.parents = ctx->sub.parents;
    p.pose = pose; p.shape = shape; p.pose_hat = pose_hat; p.shape_hat = shape_hat; p.R = R;
    p.eucl = eucl; p.eucl_pa = eucl_pa; p.angle = angle_deg; p.angle_local = angle_local != 0;
    return launch_metrics(p, static_cast<cudaStream_t>(stream));
}

int empose_metrics_joints(const float* joints, const float* joints_hat, int32_t R, float* eucl, float* eucl_pa, void* stream) {
    if (!joints || !joints_hat || !eucl || !eucl_pa || R < 0) { set_last_error("bad argument"); return EMPOSE_E_ARG; }
    MetricsParams p;
    memset(&p, 0, sizeof(p));
    p.joints = joints; p.joints_hat = joints_hat; p.R = R; p.eucl = eucl; p.eucl_pa = eucl_pa;
    return launch_metrics(p, static_cast<cudaStream_t>(stream));
}

#pragma GCC visibility pop
}
