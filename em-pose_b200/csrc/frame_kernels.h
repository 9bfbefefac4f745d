// Host interface of the per-frame CUDA kernels (frame_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fan_math.h"
#include "frame_math.h"

namespace empose {

// Input assembly: BaseModel.prepare_inputs (empose/nn/models.py:106-125) plus the per-frame weight
// of the gradient feature (loss.py:31-41 folded per frame).
struct PrepareParams {
    const float* marker_pos;    // [R][36]
    const float* marker_oris;   // [R][108]
    const int32_t* seq_len;     // [B]
    const float* masks;         // [R][12] or null
    int R, F;
    int slot_of_sensor[kSensors];   // position of each sensor in the network input, -1 if not fed
    int use_pos, use_ori, n_pos;    // n_pos = number of position columns in the network input
    int in_size;
    int in_stride, iter_stride; // row pitches of xin / xiter in elements (TMA needs 16-byte aligned rows)
    int operand_mode;           // OperandMode of xin / xiter (exact fp32, tf32-rounded fp32 or fp16 elements)
    float* meas;                // [R][12][12] exact copy, sensor-major: position (3) | orientation row-major (9)
    float* xin;                 // [R][in_stride]   (may be null)
    float* xiter;               // [R][iter_stride] columns [0, in_size) written (may be null)
    float* coef;                // [R]
};
int launch_prepare(const PrepareParams& p, cudaStream_t s);

// theta/beta update (models.py:529-535, 588-592) + feature rows [vec(R_j - I) | beta] of the blend GEMM (kFeat* in frame_math.h).
struct UpdateParams {
    float* theta;               // [R][66] in/out
    float* beta;                // [R][10] in/out
    const float* dtheta;        // [R][66] network output
    const float* dbeta;         // [R][10]
    float step;
    int first;                  // 1: theta = dtheta, beta = (mean of) dbeta  (initial estimate)
    int average_shape;
    int B, F;
    int operand_mode;           // OperandMode of xiter
    float* xiter;               // [R][iter_stride] or null: columns [in_size, in_size+76) receive theta | beta
    int in_size, iter_stride;
    float* pf;                  // [R][pf_stride] feature rows; split: [hi(kPoseFeatPad) | lo(kPoseFeatPad)]
    int pf_stride;              // kPoseFeatPad, or twice that when split
    int pf_split;               // OperandMode of the feature rows: OPERAND_F32 plain fp32; OPERAND_TF32 tf32 hi | tf32 residual;
                                // OPERAND_F16 fp16 hi | fp16 (residual * 2^11) (error-compensated blend GEMM)
    float* hist_pose;           // [R][66] or null
    float* hist_shape;          // [R][10] or null
};
int launch_update(const UpdateParams& p, cudaStream_t s);

// feature rows only (for empose_sensor_project): theta [R][66], beta [R][10]
int launch_pose_features(const float* theta, const float* beta, float* pf, int pf_stride, int pf_split, int R, cudaStream_t s);

// per-window offsets [B][12][9], [B][12][3] -> sensor-major [B][12][12] = R_off (9) | t_off (3)
int launch_pack_offsets(const float* offset_r, const float* offset_t, float* packed, int n, cudaStream_t s);

// SMPL sub-model forward (+ reverse) per frame.
struct MainParams {
    SubModel sub;
    FanModel fan;               // fan.ok: the fan-form kernel (fan_kernel.cu) runs, else the general one (frame_kernels.cu)
    ResidualSpec spec;
    const float* theta;         // [R][66]
    const float* vp;            // [R][vp_dim]  blended rest vertices v_template + S beta + P pf (the blend GEMM's output)
    const float* jrest;         // [R][kJrestLd] rest joints J0 + Jdirs beta (same GEMM)
    const float* offsets;       // [R / rows_per_offset][12][12] sensor-major R_off (9) | t_off (3), or null: use the two below
    const float* offset_r;      // [R / rows_per_offset][108]
    const float* offset_t;      // [R / rows_per_offset][36]
    int rows_per_offset;        // F for windows, 1 for per-frame offsets
    const float* meas;          // [R][12][12] sensor-major position (3) | orientation (9) (grad only)
    const float* coef;          // [R]      (grad only)
    int R;
    int want_grad;
    int round_out;              // OperandMode of dvp / dj: fp32, tf32-rounded fp32, or fp16 elements holding value * kDvpScale
    int dj_ld;                  // row pitch of dj in elements (kJrestLd, or kDjLdHalf for fp16)
    int static_tree;            // 1: sub.parents is the standard SMPL body tree -> register-resident chain phases
    long long* ticks;           // development aid: per-phase clock64() samples of one CTA, or null
    float* sensor_pos;          // [R][36] or null
    float* sensor_ori;          // [R][108] or null
    float* joints;              // [R][66] or null
    float* dvp;                 // [R][vp_dim]  (grad only) dE/dvp, NOT yet times coef
    float* dj;                  // [R][kJrestLd] (grad only) dE/dJ, NOT yet times coef (columns 66, 67 are written as zero)
    float* gtheta_part;         // [R][66]      (grad only) coef * chain part of dE/dtheta
    const float* joints_gt;     // [R][66] or null (training): adds joint_weight * d/d(theta,beta) sum_j ||J_j - Jgt_j||
    float joint_weight;
};
// development switches (empose_set_option; the environment variables EMPOSE_MAIN_GENERAL / EMPOSE_FAN_VARIANT seed them)
struct DebugOptions { int main_general; int fan_variant; int lstm_persistent; int blend_fp16; };
DebugOptions& debug_options();

int launch_main(const MainParams& p, cudaStream_t s);          // dispatches on p.fan.ok (option main_general forces the general kernel)
int launch_main_fan(const MainParams& p, cudaStream_t s);      // fan_kernel.cu

// adds the pose-blend part of dE/dtheta, scales dE/dbeta and writes the gradient features into the iter-MLP input
struct PostParams {
    const float* theta;         // [R][66]
    const float* dpf;           // [R][kPoseFeatPad] transposed blend GEMM: dE/dpf (189) | . | dE/dbeta (10 at kFeatBeta), not yet times coef
    const float* gtheta_part;   // [R][66]
    const float* coef;          // [R]
    int R;
    int operand_mode;           // OperandMode of xiter
    float* xiter;               // [R][iter_stride]: columns [in_size+76, in_size+152) receive g_theta | g_beta
    int in_size, iter_stride;
    float* g_theta_out;         // optional exact copies (tests), may be null
    float* g_beta_out;
};
int launch_post(const PostParams& p, cudaStream_t s);

// gather the last time step of a [B][F][H] sequence buffer (elements per `operand_mode`) into fp32 [B][H]
int launch_gather_last(const float* seq, float* out, int B, int F, int H, int operand_mode, cudaStream_t s);

// dst (an operand buffer of `operand_mode` elements) = src (fp32), n elements
int launch_to_operand(const float* src, float* dst, int64_t n, int operand_mode, cudaStream_t s);
// 2-D variants with row pitches (in elements): fp32 -> operand and operand -> fp32
int launch_to_operand_2d(const float* src, int64_t src_ld, int64_t rows, int cols, float* dst, int64_t dst_ld, int operand_mode,
                         cudaStream_t s);
int launch_from_operand_2d(const float* src, int64_t src_ld, int64_t rows, int cols, float* dst, int64_t dst_ld, int operand_mode,
                           cudaStream_t s);

}  // namespace empose
