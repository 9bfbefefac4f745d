// Host interface of the training-only CUDA kernels (train_kernels.cu): weight (re)packing from the flat
// parameter vector, train-mode BatchNorm1d + PReLU forward / backward, column reductions, operand transposes,
// the LSTM cell backward, the loss values and the gradient seeds of IterativeErrorFeedback.backward
// (empose/nn/models.py:634-688).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace empose {

// ---- weight packing ---------------------------------------------------------------------------------------
// dst[n][k] = src[map(n) * src_rs + k * src_cs] (+ src2[...]) for n < rows, k < cols; optional tf32 rounding.
// map(n) = n, or the LSTM gate packing of gemm_jobs.h (packed column -> torch row gate * H + unit).
struct PackOp {
    const float* src;
    const float* src2;       // optional second addend (bias_ih + bias_hh)
    float* dst;
    int32_t rows, cols;
    int64_t dst_ld;
    int64_t src_rs, src_cs;
    int32_t lstm_map;        // 1: rows are LSTM packed columns
    int32_t hidden;          // H for the LSTM map
    int32_t round;
};
int launch_pack(const PackOp* d_ops, int n_ops, int max_elems, cudaStream_t s);

// ---- BatchNorm1d (train) + PReLU ----------------------------------------------------------------------------
// z: [S*R][ld] (S segments of R rows, one per IEF iteration); columns [0, n).
// sums: double [S][2][n] (zeroed by the call): sum and sum of squares per column and segment.
int launch_col_stats(const float* z, int64_t ld, int R, int S, int n, double* sums, cudaStream_t s);
// mean / invstd [S][n] from the sums; if running_mean != null (S must be 1) update the running statistics with
// momentum (torch semantics: unbiased variance into running_var).
// `count`: rows behind every sum (R; the rows of all ranks when the sums were all-reduced for SyncBatchNorm).
int launch_bn_finalize(const double* sums, int64_t count, int S, int n, float eps, float* mean, float* invstd, float* running_mean,
                       float* running_var, float momentum, cudaStream_t s);
// a = prelu(gamma * (z - mean) * invstd + beta); gamma == null -> no BatchNorm (a = prelu(z)).
int launch_bn_apply(const float* z, int64_t ld, int R, int S, int n, const float* mean, const float* invstd,
                    const float* gamma, const float* beta, const float* alpha, int round_out, float* a, int64_t a_ld,
                    cudaStream_t s);
// backward, pass 1: per column and segment s1 = sum dy, s2 = sum dy * xhat, s3 = sum da * y * [y <= 0]
// (dy = da * prelu'(y)); sums: double [S][3][n] (zeroed by the call).
int launch_bn_bwd_reduce(const float* da, int64_t da_ld, const float* z, int64_t z_ld, int R, int S, int n, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, const float* alpha, double* sums,
                         cudaStream_t s);
// backward, pass 2: dz = gamma * invstd * (dy - s1/count - xhat * s2/count)   (no BatchNorm: dz = dy)
int launch_bn_bwd_apply(const float* da, int64_t da_ld, const float* z, int64_t z_ld, int R, int S, int n, const float* mean,
                        const float* invstd, const float* gamma, const float* beta, const float* alpha, const double* sums,
                        int64_t count, int round_out, float* dz, int64_t dz_ld, cudaStream_t s);
// parameter gradients of one BN + PReLU site from the pass-1 sums: g_gamma[c] += sum_s s2, g_beta[c] += sum_s s1,
// g_alpha += sum_{s,c} s3.  g_gamma / g_beta may be null (no BatchNorm).
int launch_bn_param_grads(const double* sums, int S, int n, float* g_gamma, float* g_beta, float* g_alpha, cudaStream_t s);

// ---- reductions / transposes --------------------------------------------------------------------------------
// dst[c] (+= if accumulate) sum over rows of src[r][c], c < n; optionally also into dst2 (bias_ih and bias_hh share it).
// scratch: double [n].
int launch_col_sum(const float* src, int64_t ld, int64_t rows, int n, double* scratch, float* dst, float* dst2, cudaStream_t s);
// dst[c][r] = src[r'][c] for c < n, r < rows, with r' = r (shift = 0) or the previous frame of the same window
// (shift = 1: r' = r - 1, zero at the first frame of every window of F frames).
// Split-K weight gradients: grad[e] += sum_i partial[i * count + e] for every table entry (fixed order: bit-reproducible)
struct DwReduce { float* grad; const float* partial; int32_t count; int32_t splits; };
int launch_dw_reduce(const DwReduce* d_table, int n_entries, int max_count, cudaStream_t s);

int launch_transpose(const float* src, int64_t src_ld, int64_t rows, int n, int shift, int F, int round_out, float* dst,
                     int64_t dst_ld, cudaStream_t s);

// ---- LSTM cell backward -------------------------------------------------------------------------------------
// One time step t of one layer for B windows: gates [B][F][4H] (activated, in the packed column order of gemm_jobs.h:
// groups of 32 = [i f g o] x 8 units), c_seq [B][F][H],
// dh_out [B][F][H] (gradient arriving at the layer output at time t), dh_rec [B][H] (from step t+1; zero at t = F-1),
// dc_rec [B][H] in/out.  Writes dgates [B][F][4H] at time t (zero rows for t >= seq_len[b]).
struct LstmCellBwdParams {
    const float* gates; const float* c_seq; const float* dh_out; const float* dh_rec; float* dc_rec; float* dgates;
    const int32_t* seq_len;
    int B, F, H, t, last, round_out;
};
int launch_lstm_cell_bwd(const LstmCellBwdParams& p, cudaStream_t s);

// ---- losses and gradient seeds ------------------------------------------------------------------------------
struct SeedParams {
    int B, F, N;                        // windows, frames, IEF iterations
    float step;
    int average_shape;
    float pose_w, shape_w, recon_w;     // loss weights (models.py:672-674); recon_w = r_weight
    int side_effect;                    // 1: the forward left d(E_i) in .grad (use_gradient, models.py:576)
    const int32_t* seq_len;             // [B]
    const float* pose_hist;             // [N+1][R][66]
    const float* shape_hist;            // [N+1][R][10]
    const float* g_theta;               // [N+1][R][66]  iterates 0..N-1: LGD feature; iterate N: final-pass gradient
    const float* g_beta;                // [N+1][R][10]    (already weighted by recon_w/(N+1) and fk_w inside the kernel)
    const float* pose_gt;               // [R][66]
    const float* shape_gt;              // [B][10]
    // outputs
    float* d_dtheta;  int64_t ld_t;     // [N][R][ld_t]  dL/d(dtheta_k)      (exact)
    float* d_dbeta;   int64_t ld_b;     // [N][R][ld_b]  dL/d(dbeta_k), pre window-mean
    float* d_init;    int64_t ld_i;     // [R][ld_i]     dL/d theta_0 in columns [0, 66), dL/d beta_0 (pre window-mean) from init_beta_col
    int init_beta_col;
    float* d_init_masked;               // [R][ld_i]     same with the rows of padded frames zeroed (or null)
};
int launch_seeds(const SeedParams& p, cudaStream_t s);

struct LossParams {
    int B, F, N;
    const int32_t* seq_len;
    const float* coef;                  // [R] = [f < len] * frame_mask * F / len  (prepare_kernel)
    const float* pose_hist; const float* shape_hist;        // [N+1][R][66|10]
    const float* markers_hist; const float* markers_ori_hist;   // [N+1][R][36|108]
    const float* joints_final;          // [R][66]
    const float* meas;                  // [R][12][12] sensor-major: position (3) | orientation (9)
    const float* pose_gt; const float* shape_gt; const float* joints_gt;   // joints_gt may be null
    int sensor_active[12];
    int use_pos, use_ori;
    double* sums;                       // [4]: pose, shape, reconstruction, fk  (sums over iterates of the batch means)
};
int launch_losses(const LossParams& p, cudaStream_t s);

// x[r][c] = round_tf32(x[r][c]) for c < n (row pitch ld)
int launch_round_inplace(float* x, int64_t rows, int n, int64_t ld, cudaStream_t s);
// dst[i] += src[i] for n floats (n % 4 == 0, 16-byte aligned), optionally rounded to tf32 (a LinearLayers skip, layers.py:35-43)
int launch_add_inplace(float* dst, const float* src, int64_t n, int round_out, cudaStream_t s);

}  // namespace empose
