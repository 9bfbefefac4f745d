// Full-mesh SMPL-H evaluation behind SMPLLayer.forward (reference empose/bodymodels/smpl.py:81-165, i.e. the
// third-party BodyModel call at smpl.py:121 with a zero hand pose).
//
//   pose kernel   per frame: Rodrigues, J = J0 + Jdirs beta (52 joints), 22-joint chain -> skinning transforms A,
//                 posed joints (hand joints ride on their wrist), pose features
//   pose blend    [N x 189] . [189 x 3V] on the GEMM job executor (tcgen05 in TF32 mode, error-compensated 3xTF32)
//   skin kernel   per (frame, vertex): v_template + S beta + pose blend, linear blend skinning, + trans
//
// The output (82.7 KB per frame) makes this HBM-bound; frames are processed in slabs so the pose-blend scratch stays
// bounded.
#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <vector>

#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_math.h"
#include "gemm_jobs.h"
#include "gemm_tc.h"
#include "model_internal.h"

namespace empose {

// K of the full-mesh pose-blend GEMM (189 pose features padded to a multiple of 32 floats); the sub-model path has its own,
// wider feature row (frame_math.h kPoseFeatPad) because its GEMM also carries the shape blend.
constexpr int kFullFeatPad = 192;
namespace {

constexpr int kAllJoints = 52;
constexpr int kSlabFrames = 2048;

struct FullModel {
    int n_verts = 0, n_skin = 0, v3 = 0;
    float *v_template = nullptr, *shapedirs = nullptr, *j0 = nullptr, *jdirs = nullptr, *skin_weight = nullptr;
    int *parents = nullptr, *ancestor = nullptr, *skin_joint = nullptr;
};

// one warp per frame
__global__ void __launch_bounds__(128) smpl_pose_kernel(FullModel fm, const float* __restrict__ poses_root,
                                                        const float* __restrict__ poses_body, const float* __restrict__ betas,
                                                        const float* __restrict__ trans, int n, int pf_stride, int pf_split,
                                                        float* __restrict__ pf, float* __restrict__ amat, float* __restrict__ joints) {
    __shared__ float s_rot[4][kJoints][9], s_j[4][kAllJoints][3], s_g[4][kJoints][12], s_theta[4][kPoseDim], s_beta[4][kBetas];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * 4 + warp;
    if (f >= n) return;
    for (int i = lane; i < kPoseDim; i += 32)
        s_theta[warp][i] = i < 3 ? (poses_root ? poses_root[f * 3 + i] : 0.0f) : poses_body[f * 63 + (i - 3)];
    if (lane < kBetas) s_beta[warp][lane] = betas[f * kBetas + lane];
    __syncwarp();
    for (int j = lane; j < kJoints; j += 32) {
        rodrigues_fwd(&s_theta[warp][j * 3], s_rot[warp][j]);
        if (j > 0) {
            float* dst = pf + f * pf_stride + (j - 1) * 9;
            for (int e = 0; e < 9; ++e) {
                const float v = s_rot[warp][j][e] - ((e % 4 == 0) ? 1.0f : 0.0f);
                if (pf_split) { const float hi = round_tf32(v); dst[e] = hi; dst[kFullFeatPad + e] = round_tf32(v - hi); }
                else dst[e] = v;
            }
        }
    }
    for (int i = lane; i < kAllJoints * 3; i += 32) {
        float acc = fm.j0[i];
        for (int k = 0; k < kBetas; ++k) acc += fm.jdirs[k * kAllJoints * 3 + i] * s_beta[warp][k];
        s_j[warp][i / 3][i % 3] = acc;
    }
    __syncwarp();
    if (lane < 3) {                                  // row-parallel chain, as in frame_math.h phase_chain
        const int r = lane;
        float (*g)[12] = s_g[warp];
        for (int c = 0; c < 3; ++c) g[0][r * 3 + c] = s_rot[warp][0][r * 3 + c];
        g[0][9 + r] = s_j[warp][0][r];
        for (int j = 1; j < kJoints; ++j) {
            const int p = fm.parents[j];
            const float g0 = g[p][r * 3], g1 = g[p][r * 3 + 1], g2 = g[p][r * 3 + 2];
            const float* R = s_rot[warp][j];
            g[j][r * 3 + 0] = g0 * R[0] + g1 * R[3] + g2 * R[6];
            g[j][r * 3 + 1] = g0 * R[1] + g1 * R[4] + g2 * R[7];
            g[j][r * 3 + 2] = g0 * R[2] + g1 * R[5] + g2 * R[8];
            g[j][9 + r] = g0 * (s_j[warp][j][0] - s_j[warp][p][0]) + g1 * (s_j[warp][j][1] - s_j[warp][p][1]) +
                          g2 * (s_j[warp][j][2] - s_j[warp][p][2]) + g[p][9 + r];
        }
    }
    __syncwarp();
    const float t0 = trans ? trans[f * 3] : 0.0f, t1 = trans ? trans[f * 3 + 1] : 0.0f, t2 = trans ? trans[f * 3 + 2] : 0.0f;
    const float tr[3] = {t0, t1, t2};
    if (joints)
        for (int i = lane; i < kAllJoints * 3; i += 32) {
            const int j = i / 3, r = i % 3, a = fm.ancestor[j];
            const float* g = s_g[warp][a];
            float v = g[9 + r];
            if (j >= kJoints)                        // zero-pose hand joint: rigidly attached to its body ancestor
                v += g[r * 3] * (s_j[warp][j][0] - s_j[warp][a][0]) + g[r * 3 + 1] * (s_j[warp][j][1] - s_j[warp][a][1]) +
                     g[r * 3 + 2] * (s_j[warp][j][2] - s_j[warp][a][2]);
            joints[f * kAllJoints * 3 + i] = v + tr[r];
        }
    // skinning transforms A_j = [G^R | G^t - G^R J_j + trans]
    for (int i = lane; i < kJoints * 12; i += 32) {
        const int j = i / 12, e = i % 12;
        const float* g = s_g[warp][j];
        float v;
        if (e < 9) v = g[e];
        else {
            const int r = e - 9;
            v = g[9 + r] - (g[r * 3] * s_j[warp][j][0] + g[r * 3 + 1] * s_j[warp][j][1] + g[r * 3 + 2] * s_j[warp][j][2]) + tr[r];
        }
        amat[f * kJoints * 12 + i] = v;
    }
}

// thread per (frame, vertex); consecutive threads = consecutive vertices of one frame
__global__ void __launch_bounds__(256) smpl_skin_kernel(FullModel fm, const float* __restrict__ betas, const float* __restrict__ vp_off,
                                                        int64_t vp_stride, const float* __restrict__ amat, int n,
                                                        float* __restrict__ verts) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * fm.n_verts) return;
    const int64_t f = idx / fm.n_verts;
    const int v = (int)(idx % fm.n_verts);
    float b[kBetas];
#pragma unroll
    for (int k = 0; k < kBetas; ++k) b[k] = betas[f * kBetas + k];
    float p[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = fm.v_template[v * 3 + c] + vp_off[f * vp_stride + v * 3 + c];
#pragma unroll
        for (int k = 0; k < kBetas; ++k) acc += fm.shapedirs[(int64_t)k * fm.v3 + v * 3 + c] * b[k];
        p[c] = acc;
    }
    float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;
    for (int s = 0; s < fm.n_skin; ++s) {
        const float w = fm.skin_weight[v * fm.n_skin + s];
        const float* A = amat + (f * kJoints + fm.skin_joint[v * fm.n_skin + s]) * 12;
        x0 += w * (A[0] * p[0] + A[1] * p[1] + A[2] * p[2] + A[9]);
        x1 += w * (A[3] * p[0] + A[4] * p[1] + A[5] * p[2] + A[10]);
        x2 += w * (A[6] * p[0] + A[7] * p[1] + A[8] * p[2] + A[11]);
    }
    float* o = verts + idx * 3;
    o[0] = x0; o[1] = x1; o[2] = x2;
}

}  // namespace
}  // namespace empose

using namespace empose;

struct empose_smpl {
    int device = 0, num_sms = 148;
    bool round = true;
    int pf_stride = kFullFeatPad;
    Arena arena;
    FullModel fm;
    PackedMatrix pb;
    // slab workspace + jobs
    float *pf = nullptr, *vp_off = nullptr, *amat = nullptr;
    JobBook book;
    JobRange range;
};

extern "C" {
#pragma GCC visibility push(default)

int empose_smpl_create(const empose_tensor* tensors, int32_t n_tensors, int32_t precision, int32_t device, empose_smpl** out) {
    if (!tensors || !out) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device available: empose_b200 has no CPU fallback");
        return EMPOSE_E_CUDA;
    }
    EMPOSE_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    EMPOSE_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_last_error("empose_b200 is built for sm_100a (B200) only"); return EMPOSE_E_CUDA; }
    std::unique_ptr<empose_smpl> ctx(new empose_smpl());
    ctx->device = device; ctx->num_sms = prop.multiProcessorCount;
    ctx->round = precision != EMPOSE_PRECISION_FP32;      // TF32 and FP16 modes both run the pose blend as error-compensated tf32
    ctx->pf_stride = ctx->round ? 2 * kFullFeatPad : kFullFeatPad;
    TensorTable tt{tensors, n_tensors};
    const int32_t* dims;
    EMPOSE_TRY(tt.get_i32("smpl.dims", 3, &dims));
    FullModel& fm = ctx->fm;
    fm.n_verts = dims[0]; fm.n_skin = dims[2]; fm.v3 = fm.n_verts * 3;
    if (dims[1] != kAllJoints || fm.n_verts < 1 || fm.n_skin < 1 || fm.n_skin > kJoints) { set_last_error("smpl.dims out of range"); return EMPOSE_E_ARG; }
    Arena& A = ctx->arena;
    auto up_f = [&](const char* name, int64_t n, float** dst) -> int {
        const float* h;
        EMPOSE_TRY(tt.get_f32(name, {n}, &h));
        return A.upload(std::vector<float>(h, h + n), dst);
    };
    auto up_i = [&](const char* name, int64_t n, int lo, int hi, int** dst) -> int {
        const int32_t* h;
        EMPOSE_TRY(tt.get_i32(name, n, &h));
        for (int64_t i = 0; i < n; ++i)
            if (h[i] < lo || h[i] >= hi) { set_last_error(std::string("index out of range in '") + name + "'"); return EMPOSE_E_ARG; }
        return A.upload(std::vector<int>(h, h + n), dst);
    };
    EMPOSE_TRY(up_f("smpl.v_template", fm.v3, &fm.v_template));
    EMPOSE_TRY(up_f("smpl.shapedirs", (int64_t)kBetas * fm.v3, &fm.shapedirs));
    EMPOSE_TRY(up_f("smpl.j0", kAllJoints * 3, &fm.j0));
    EMPOSE_TRY(up_f("smpl.jdirs", (int64_t)kBetas * kAllJoints * 3, &fm.jdirs));
    EMPOSE_TRY(up_f("smpl.skin_weight", (int64_t)fm.n_verts * fm.n_skin, &fm.skin_weight));
    EMPOSE_TRY(up_i("smpl.skin_joint", (int64_t)fm.n_verts * fm.n_skin, 0, kJoints, &fm.skin_joint));
    EMPOSE_TRY(up_i("smpl.parents", kJoints, -1, kJoints, &fm.parents));
    EMPOSE_TRY(up_i("smpl.ancestor", kAllJoints, 0, kJoints, &fm.ancestor));
    // pose-blend matrix W[i][k] = P[k][i], N = 3V, error-compensated in TF32 mode (see model.cu upload_submodel)
    const float* P;
    EMPOSE_TRY(tt.get_f32("smpl.posedirs", {kPoseFeat, fm.v3}, &P));
    const int v3 = fm.v3;
    if (ctx->round) {
        std::vector<float> w0((size_t)v3 * 2 * kFullFeatPad, 0.0f), w1((size_t)v3 * kPoseFeat, 0.0f);
        for (int i = 0; i < v3; ++i)
            for (int k = 0; k < kPoseFeat; ++k) {
                const float v = P[(size_t)k * v3 + i];
                const float hi = host_round_tf32(v);
                w0[(size_t)i * 2 * kFullFeatPad + k] = hi;
                w0[(size_t)i * 2 * kFullFeatPad + kFullFeatPad + k] = hi;
                w1[(size_t)i * kPoseFeat + k] = host_round_tf32(v - hi);
            }
        EMPOSE_TRY(pack_matrix(A, v3, 2 * kFullFeatPad, kPoseFeat, 16, true, false, [&](int r) {
            return RowSource{&w0[(size_t)r * 2 * kFullFeatPad], &w1[(size_t)r * kPoseFeat], 1.0, 0.0};
        }, &ctx->pb));
    } else {
        std::vector<float> pt((size_t)v3 * kPoseFeat);
        for (int i = 0; i < v3; ++i)
            for (int k = 0; k < kPoseFeat; ++k) pt[(size_t)i * kPoseFeat + k] = P[(size_t)k * v3 + i];
        EMPOSE_TRY(pack_matrix(A, v3, kPoseFeat, 0, 16, false, false,
                               [&](int r) { return RowSource{&pt[(size_t)r * kPoseFeat], nullptr, 1.0, 0.0}; }, &ctx->pb));
    }
    // slab workspace and the pose-blend jobs over it
    const int64_t vp_stride = ctx->pb.n_pad;
    EMPOSE_TRY(A.alloc_n((size_t)kSlabFrames * ctx->pf_stride, &ctx->pf, true));
    EMPOSE_TRY(A.alloc_n((size_t)kSlabFrames * vp_stride, &ctx->vp_off));
    EMPOSE_TRY(A.alloc_n((size_t)kSlabFrames * kJoints * 12, &ctx->amat));
    ctx->book.use_tc = ctx->round;
    GemmJob proto = linear_proto(ctx->pb, false, ctx->vp_off, vp_stride, v3);
    ASrc a0{ctx->pf, ctx->pf_stride, ctx->round ? 2 * kFullFeatPad : kFullFeatPad, kSlabFrames};
    ASrc a1 = ctx->round ? ASrc{ctx->pf, ctx->pf_stride, kFullFeatPad, kSlabFrames} : ASrc{};
    EMPOSE_TRY(ctx->book.add(ctx->pb, a0, a1, proto, kSlabFrames, -1, &ctx->range));
    EMPOSE_TRY(ctx->book.finalize(A));
    *out = ctx.release();
    return EMPOSE_OK;
}

void empose_smpl_destroy(empose_smpl* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    delete ctx;
}

int empose_smpl_forward(empose_smpl* ctx, const float* poses_root, const float* poses_body, const float* betas,
                        const float* trans, int32_t N, float* verts, float* joints, void* stream) {
    if (!ctx || !poses_body || !betas || N < 1) { set_last_error("bad argument"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const FullModel& fm = ctx->fm;
    const int64_t vp_stride = ctx->pb.n_pad;
    for (int begin = 0; begin < N; begin += kSlabFrames) {
        const int n = N - begin < kSlabFrames ? N - begin : kSlabFrames;
        const float* pr = poses_root ? poses_root + (size_t)begin * 3 : nullptr;
        const float* tr = trans ? trans + (size_t)begin * 3 : nullptr;
        smpl_pose_kernel<<<ceil_div(n, 4), 128, 0, s>>>(fm, pr, poses_body + (size_t)begin * 63, betas + (size_t)begin * kBetas, tr, n,
                                                        ctx->pf_stride, ctx->round ? 1 : 0, ctx->pf, ctx->amat,
                                                        joints ? joints + (size_t)begin * kAllJoints * 3 : nullptr);
        EMPOSE_CUDA_TRY(cudaGetLastError());
        if (!verts) continue;
        // rows beyond n in the slab hold stale features; their products are never read
        if (ctx->round) EMPOSE_TRY(tc_launch(ctx->book.d_jobs, ctx->book.d_maps, ctx->range.begin, ctx->range.count, 1, ceil_div(n, kTileM), ctx->num_sms, s));
        else EMPOSE_TRY(simt_launch(ctx->book.d_jobs, ctx->book.jobs.data(), ctx->range.begin, ctx->range.count, ceil_div(n, kTileM), s, nullptr));
        const int64_t total = (int64_t)n * fm.n_verts;
        smpl_skin_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(fm, betas + (size_t)begin * kBetas, ctx->vp_off, vp_stride, ctx->amat, n,
                                                                      verts + (size_t)begin * fm.n_verts * 3);
        EMPOSE_CUDA_TRY(cudaGetLastError());
    }
    return EMPOSE_OK;
}

#pragma GCC visibility pop
}  // extern "C"
