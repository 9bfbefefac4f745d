// Host interface of the per-frame CUDA kernels (frame_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

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
    float* meas;                // [R][144] exact copy [pos | ori]
    float* xin;                 // [R][in_stride]   (may be null)
    float* xiter;               // [R][iter_stride] columns [0, in_size) written (may be null)
    float* coef;                // [R]
};
int launch_prepare(const PrepareParams& p, cudaStream_t s);

// theta/beta update (models.py:529-535, 588-592) + pose features for the pose-blend GEMM.
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
    float* pf;                  // [R][pf_stride] pose features vec(R_1..R_21 - I); split: [hi(192) | lo(192)]
    int pf_stride;              // 192, or 384 when split
    int pf_split;               // 1: write tf32 hi part and tf32 residual (error-compensated pose-blend GEMM)
    float* hist_pose;           // [R][66] or null
    float* hist_shape;          // [R][10] or null
};
int launch_update(const UpdateParams& p, cudaStream_t s);

// pose features only (for empose_sensor_project)
int launch_pose_features(const float* theta, float* pf, int pf_stride, int pf_split, int R, cudaStream_t s);

// SMPL sub-model forward (+ reverse) per frame.
struct MainParams {
    SubModel sub;
    ResidualSpec spec;
    const float* theta;         // [R][66]
    const float* beta;          // [R][10]
    const float* vp_off;        // [R][vp_dim]  pose-blend result
    const float* offset_r;      // [R / rows_per_offset][108]
    const float* offset_t;      // [R / rows_per_offset][36]
    int rows_per_offset;        // F for windows, 1 for per-frame offsets
    const float* meas;          // [R][144] (grad only)
    const float* coef;          // [R]      (grad only)
    int R;
    int want_grad;
    int round_out;
    int static_tree;            // 1: sub.parents is the standard SMPL body tree -> register-resident chain phases
    int legacy_blend;           // 1: per-item shape-blend phases of frame_math.h instead of the CTA-wide vector ones (set by launch_main
                                // from EMPOSE_MAIN_LEGACY_BLEND; A/B measurements)
    long long* ticks;           // development aid: per-phase clock64() samples of one CTA, or null
    float* sensor_pos;          // [R][36] or null
    float* sensor_ori;          // [R][108] or null
    float* joints;              // [R][66] or null
    float* dvp;                 // [R][vp_dim]  (grad only)
    float* gtheta_part;         // [R][66]      (grad only) coef * chain part of dE/dtheta
    float* gbeta;               // [R][10]      (grad only) coef * dE/dbeta
    const float* joints_gt;     // [R][66] or null (training): adds joint_weight * d/d(theta,beta) sum_j ||J_j - Jgt_j||
    float joint_weight;
};
int launch_main(const MainParams& p, cudaStream_t s);

// adds the pose-blend part of dE/dtheta and writes the gradient features into the iter-MLP input
struct PostParams {
    const float* theta;         // [R][66]
    const float* dpf;           // [R][192]
    const float* gtheta_part;   // [R][66]
    const float* gbeta;         // [R][10]
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
