// The (Bi)RNN baseline of the reference (SimpleRNN, empose/nn/models.py:265-317) on the same job executors as the LGD
// path (C ABI: empose_rnn_*, include/empose_b200.h).
//
//   prepare_inputs (models.py:106-125)
//   -> L layers of a uni- or bidirectional LSTM with packed-sequence semantics (layers.py:133-157).  Every time step of
//      a layer is one launch holding the step of BOTH directions (forward at t = s, reverse at t = F-1-s); a layer's
//      output [h_forward | h_reverse] is one [B][F][dirs*H] buffer, which is at once the next layer's input, the
//      recurrent operand of the next step and -- for rows whose sequence has not started / has ended -- the carried
//      state (the reverse direction of a padded sequence simply keeps its initial state until t = len-1).
//   -> to_pose (models.py:298) and, with m_estimate_shape, the BatchNorm-free to_shape MLP (models.py:276-280, 302-306)
//      chained per 128-row tile inside one CTA, the per-window shape mean, and maybe_do_fk (models.py:134-144) through
//      the SMPL sub-model kernels of the LGD path.
// Inference only; m_learn_init_state is not supported (see oracle/rnn.py for why the reference's version cannot be
// combined with a bidirectional LSTM).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_kernels.h"
#include "gemm_jobs.h"
#include "gemm_tc.h"
#include "model_internal.h"
#include "rnn_persistent.h"

namespace empose {
namespace {

struct RnnPlan {
    int B = 0, F = 0, R = 0;
    Arena arena;
    JobBook book;
    float *meas = nullptr, *xin = nullptr, *coef = nullptr;
    int32_t* seq_len = nullptr;
    std::vector<float*> hseq;                       // [layer]: [B][F][dirs*H] operand elements
    std::vector<float*> hinit, cstate;              // [layer*dirs + dir]: [B][H]
    float *pose = nullptr, *dshape = nullptr, *theta = nullptr, *beta = nullptr, *pf = nullptr, *vpoff = nullptr, *joints = nullptr;
    float *jrest = nullptr, *offsets = nullptr;
    float* act[2] = {nullptr, nullptr};
    int64_t act_rows = 0;
    std::vector<JobRange> steps;                    // [layer * F + s]
    JobRange to_pose, to_shape, pb;
    // persistent single-sequence path (B == 1, fp16 mode): input projections as one GEMM per layer, W_hh resident in smem
    bool persistent = false;
    int pC = 0, pU = 0;
    float* xw[2] = {nullptr, nullptr};              // [F][4H] per direction, reused by every layer
    float* hx = nullptr;
    unsigned* counters = nullptr;
    std::vector<JobRange> xw_jobs;                  // [layer]
};

__global__ void gather_state_kernel(const float* __restrict__ seq, int64_t pitch, int t, int col0, float* __restrict__ out, int B,
                                    int F, int H, int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H) return;
    const int64_t b = i / H;
    const int u = (int)(i % H);
    out[i] = load_operand(seq, (b * F + t) * pitch + col0 + u, mode);
}

// one row of sensor-major offsets [12][12] = R_off (9) | t_off (3): identity, zero
__global__ void identity_offsets_kernel(float* __restrict__ offsets) {
    const int i = threadIdx.x;
    if (i < 144) offsets[i] = ((i % 12) < 9 && ((i % 12) % 4) == 0) ? 1.0f : 0.0f;
}

}  // namespace
}  // namespace empose

using namespace empose;

struct empose_rnn {
    empose_rnn_config cfg;
    IefData fk;                   // sub-model, pose-blend operands, arena, executor flags (weights of the LGD model unused)
    int in_size = 0, in_stride = 0, n_pos = 0, dirs = 1;
    int slot_of_sensor[kSensors];
    std::vector<PackedMatrix> lstm;      // [layer * dirs + dir]
    std::vector<PackedMatrix> wih_plain; // fp16 mode: W_ih alone, torch row order, bias = b_ih + b_hh (persistent path)
    std::vector<__half*> whh_half;       // fp16 mode: W_hh [4H][H] fp16, torch row order
    PackedMatrix to_pose;
    MlpPacked to_shape;
    std::unique_ptr<RnnPlan> plan;
    int64_t launches = 0;
};

namespace empose {
namespace {

int run(empose_rnn* ctx, RnnPlan& pl, const JobRange& r, int m_tiles, cudaStream_t s) {
    if (r.count == 0) return EMPOSE_OK;
    if (ctx->fk.round) {
        ++ctx->launches;
        return tc_launch(pl.book.d_jobs, pl.book.d_maps, r.begin, r.count, r.per_item, m_tiles, ctx->fk.num_sms, s);
    }
    return simt_launch(pl.book.d_jobs, pl.book.jobs.data(), r.begin, r.count, m_tiles, s, &ctx->launches);
}

int build_plan(empose_rnn* ctx, int B, int F, RnnPlan** out) {
    if (ctx->plan && ctx->plan->B == B && ctx->plan->F == F) { *out = ctx->plan.get(); return EMPOSE_OK; }
    ctx->plan.reset();
    std::unique_ptr<RnnPlan> plp(new RnnPlan());
    RnnPlan& pl = *plp;
    const empose_rnn_config& cfg = ctx->cfg;
    const IefData& fk = ctx->fk;
    pl.B = B; pl.F = F; pl.R = B * F;
    const int R = pl.R, H = cfg.hidden_size, L = cfg.num_layers, D = ctx->dirs, W = D * H;
    const int hf = fk.op_half;
    const size_t esz = hf ? 2 : 4, Rz = (size_t)R;
    pl.book.use_tc = fk.round;
    Arena& A = pl.arena;
    auto alloc_operand = [&](size_t elems, float** o) -> int {
        void* p;
        EMPOSE_TRY(A.alloc(elems * esz, &p, true));
        *o = static_cast<float*>(p);
        return EMPOSE_OK;
    };
    EMPOSE_TRY(A.alloc_n(Rz * 144, &pl.meas));
    EMPOSE_TRY(A.alloc_n(Rz, &pl.coef));
    EMPOSE_TRY(alloc_operand(Rz * ctx->in_stride, &pl.xin));
    EMPOSE_TRY(A.alloc_n((size_t)B, &pl.seq_len));
    EMPOSE_TRY(A.alloc_n(Rz * kPoseDim, &pl.pose));
    pl.hseq.resize(L); pl.hinit.resize(L * D); pl.cstate.resize(L * D);
    for (int l = 0; l < L; ++l) {
        EMPOSE_TRY(alloc_operand(Rz * W, &pl.hseq[l]));
        for (int d = 0; d < D; ++d) {
            EMPOSE_TRY(alloc_operand((size_t)B * H, &pl.hinit[l * D + d]));
            EMPOSE_TRY(A.alloc_n((size_t)B * H, &pl.cstate[l * D + d], true));
        }
    }
    // ---- single stream: persistent recurrence (rnn_persistent.cu) instead of one launch per time step ----
    pl.persistent = B == 1 && hf && !ctx->whh_half.empty() && F >= 64 && getenv("EMPOSE_RNN_NO_PERSISTENT") == nullptr &&
                    lstm_persistent_pick(H, D, fk.num_sms, &pl.pC, &pl.pU);
    if (pl.persistent) {
        for (int d = 0; d < D; ++d) EMPOSE_TRY(A.alloc_n(Rz * 4 * H, &pl.xw[d]));
        EMPOSE_TRY(A.alloc_n((size_t)D * 2 * H, &pl.hx, true));
        EMPOSE_TRY(A.alloc_n((size_t)64, &pl.counters, true));
        pl.xw_jobs.resize(L);
        for (int l = 0; l < L; ++l)
            for (int d = 0; d < D; ++d) {
                const PackedMatrix& Wm = ctx->wih_plain[(size_t)l * D + d];
                ASrc a0 = l == 0 ? ASrc{pl.xin, ctx->in_stride, ctx->in_size, R, hf} : ASrc{pl.hseq[l - 1], W, W, R, hf};
                GemmJob proto = linear_proto(Wm, false, pl.xw[d], 4 * H, 4 * H);
                EMPOSE_TRY(pl.book.add(Wm, a0, ASrc{}, proto, R, -1, &pl.xw_jobs[l]));
            }
    }
    // ---- LSTM steps ----
    pl.steps.resize(pl.persistent ? 0 : (size_t)L * F);
    for (int l = 0; l < L && !pl.persistent; ++l)
        for (int s = 0; s < F; ++s)
            for (int d = 0; d < D; ++d) {
                const int t = d == 0 ? s : F - 1 - s;
                const int t_prev = d == 0 ? t - 1 : t + 1;
                ASrc a0 = s == 0 ? ASrc{pl.hinit[l * D + d], H, H, B, hf}
                                 : ASrc{operand_at(pl.hseq[l], (size_t)t_prev * W + (size_t)d * H, hf), (int64_t)F * W, H, B, hf};
                ASrc a1 = l == 0 ? ASrc{operand_at(pl.xin, (size_t)t * ctx->in_stride, hf), (int64_t)F * ctx->in_stride, ctx->in_size, B, hf}
                                 : ASrc{operand_at(pl.hseq[l - 1], (size_t)t * W, hf), (int64_t)F * W, W, B, hf};
                GemmJob proto;
                memset(&proto, 0, sizeof(proto));
                proto.epi = EPI_LSTM;
                proto.round_out = fk.round ? 1 : 0;
                proto.out_half = hf;
                proto.out = operand_at(pl.hseq[l], (size_t)t * W + (size_t)d * H, hf);
                proto.out_stride = (int64_t)F * W;
                proto.c_state = pl.cstate[l * D + d];
                proto.h_prev = a0.ptr;
                proto.h_prev_stride = a0.stride;
                proto.t = t;
                proto.hidden = H;
                proto.seq_len = pl.seq_len;
                proto.frames_per_window = 1;
                proto.split = 1 << 30;
                EMPOSE_TRY(pl.book.add(ctx->lstm[l * D + d], a0, a1, proto, B, -1, &pl.steps[(size_t)l * F + s]));
            }
    // ---- to_pose: rows of padded frames see a zero LSTM output (pad_packed_sequence) -> bias only ----
    {
        GemmJob proto = linear_proto(ctx->to_pose, false, pl.pose, kPoseDim, kPoseDim);
        proto.mask_rows = R; proto.seq_len = pl.seq_len; proto.frames_per_window = F;
        EMPOSE_TRY(pl.book.add(ctx->to_pose, ASrc{pl.hseq[L - 1], W, W, R, hf}, ASrc{}, proto, R, -1, &pl.to_pose));
    }
    if (cfg.estimate_shape) {
        const int SH = cfg.shape_hidden_size;
        EMPOSE_TRY(A.alloc_n(Rz * kBetas, &pl.dshape));
        EMPOSE_TRY(A.alloc_n(Rz * kPoseDim, &pl.theta));
        EMPOSE_TRY(A.alloc_n(Rz * kBetas, &pl.beta));
        EMPOSE_TRY(A.alloc_n(Rz * fk.pf_stride, &pl.pf, true));
        const bool scratch = fk.round;
        pl.act_rows = scratch ? (int64_t)fk.num_sms * kTileM : (int64_t)R;
        for (int q = 0; q < 2; ++q) EMPOSE_TRY(alloc_operand((size_t)pl.act_rows * SH, &pl.act[q]));
        const int nl = (int)ctx->to_shape.layers.size();
        int last = -1;
        for (int l = 0; l < nl; ++l) {
            const PackedMatrix& Wm = ctx->to_shape.layers[l];
            ASrc a0 = l == 0 ? ASrc{pl.hseq[L - 1], W, W, R, hf} : ASrc{pl.act[(l - 1) & 1], SH, SH, pl.act_rows, hf};
            GemmJob proto;
            int m_rows = R;
            if (l == nl - 1) {
                proto = linear_proto(Wm, false, pl.dshape, kBetas, kBetas);
            } else {
                proto = linear_proto(Wm, fk.op_mode == OPERAND_TF32, pl.act[l & 1], SH, SH);
                proto.out_half = hf;
                if (scratch) { proto.out_scratch = 1; m_rows = (int)pl.act_rows; }
            }
            if (l == 0) { proto.mask_rows = R; proto.seq_len = pl.seq_len; proto.frames_per_window = F; }
            if (scratch && l > 0) proto.a_scratch[0] = 1;
            EMPOSE_TRY(pl.book.add(Wm, a0, ASrc{}, proto, m_rows, last, &pl.to_shape));
            last = pl.to_shape.count - 1;
        }
        pl.to_shape.per_item = pl.to_shape.count;
        if (cfg.do_fk) {
            const int vp = fk.sub.vp_dim;
            EMPOSE_TRY(A.alloc_n(Rz * vp, &pl.vpoff));
            EMPOSE_TRY(A.alloc_n(Rz * kPoseDim, &pl.joints));
            EMPOSE_TRY(A.alloc_n((size_t)R * kJrestLd, &pl.jrest));
            EMPOSE_TRY(A.alloc_n((size_t)144, &pl.offsets));
            EMPOSE_TRY(add_blend_jobs(pl.book, &fk, pl.pf, pl.vpoff, pl.jrest, R, &pl.pb));
        }
    }
    EMPOSE_TRY(pl.book.finalize(A));
    *out = plp.get();
    ctx->plan = std::move(plp);
    return EMPOSE_OK;
}

int rnn_forward(empose_rnn* ctx, RnnPlan& pl, const float* marker_pos, const float* marker_oris, const int32_t* seq_lengths,
                float* lstm_state, int is_new_sequence, float* pose_hat, float* shape_hat, float* joints_hat, cudaStream_t s) {
    const empose_rnn_config& cfg = ctx->cfg;
    const IefData& fk = ctx->fk;
    const int B = pl.B, F = pl.F, R = pl.R, H = cfg.hidden_size, L = cfg.num_layers, D = ctx->dirs, W = D * H;
    const int mt_R = ceil_div(R, kTileM), mt_B = ceil_div(B, kTileM);
    ctx->launches = 0;
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.seq_len, seq_lengths, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    PrepareParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.marker_pos = marker_pos; pp.marker_oris = marker_oris; pp.seq_len = pl.seq_len; pp.masks = nullptr;
    pp.R = R; pp.F = F;
    for (int i = 0; i < kSensors; ++i) pp.slot_of_sensor[i] = ctx->slot_of_sensor[i];
    pp.use_pos = cfg.use_marker_pos; pp.use_ori = cfg.use_marker_ori; pp.n_pos = ctx->n_pos;
    pp.in_size = ctx->in_size; pp.in_stride = ctx->in_stride; pp.iter_stride = 0; pp.operand_mode = fk.op_mode;
    pp.meas = pl.meas; pp.xin = pl.xin; pp.xiter = nullptr; pp.coef = pl.coef;
    EMPOSE_TRY(launch_prepare(pp, s));
    ++ctx->launches;
    const size_t st_bytes = (size_t)B * H * 4;
    const int LD = L * D;
    for (int q = 0; q < LD; ++q) {
        if (lstm_state && !is_new_sequence) {
            EMPOSE_TRY(launch_to_operand(lstm_state + (size_t)q * B * H, pl.hinit[q], (int64_t)B * H, fk.op_mode, s));
            EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.cstate[q], lstm_state + (size_t)(LD + q) * B * H, st_bytes, cudaMemcpyDeviceToDevice, s));
            ++ctx->launches;
        } else {
            EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.hinit[q], 0, (size_t)B * H * operand_bytes(fk.op_mode), s));
            EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.cstate[q], 0, st_bytes, s));
        }
    }
    if (pl.persistent) {
        int len_host = F;          // B == 1: the one sequence length (device -> host: the kernel takes it as a scalar)
        EMPOSE_CUDA_TRY(cudaMemcpyAsync(&len_host, pl.seq_len, sizeof(int), cudaMemcpyDeviceToHost, s));
        EMPOSE_CUDA_TRY(cudaStreamSynchronize(s));
        for (int l = 0; l < L; ++l) {
            EMPOSE_TRY(run(ctx, pl, pl.xw_jobs[l], mt_R, s));
            LstmPersistentParams lp;
            memset(&lp, 0, sizeof(lp));
            for (int d = 0; d < D; ++d) { lp.w_hh[d] = ctx->whh_half[(size_t)l * D + d]; lp.xw[d] = pl.xw[d]; }
            const bool carry = lstm_state && !is_new_sequence;
            lp.h0 = carry ? lstm_state + (size_t)l * D * H : nullptr;
            lp.c0 = carry ? lstm_state + (size_t)(LD + l * D) * H : nullptr;
            lp.h_out = lstm_state ? lstm_state + (size_t)l * D * H : nullptr;
            lp.c_out = lstm_state ? lstm_state + (size_t)(LD + l * D) * H : nullptr;
            lp.hseq = pl.hseq[l]; lp.hseq_pitch = W; lp.hseq_mode = fk.op_mode;
            lp.hx = pl.hx; lp.counters = pl.counters;
            lp.F = F; lp.H = H; lp.dirs = D; lp.C = pl.pC; lp.U = pl.pU;
            lp.len = len_host < F ? len_host : F;
            EMPOSE_TRY(launch_lstm_persistent(lp, s));
            ++ctx->launches;
        }
    }
    for (const JobRange& r : pl.steps) EMPOSE_TRY(run(ctx, pl, r, mt_B, s));
    if (lstm_state && !pl.persistent)
        for (int l = 0; l < L; ++l)
            for (int d = 0; d < D; ++d) {
                const int q = l * D + d;
                gather_state_kernel<<<(unsigned)(((int64_t)B * H + 255) / 256), 256, 0, s>>>(pl.hseq[l], W, d == 0 ? F - 1 : 0, d * H,
                                                                                              lstm_state + (size_t)q * B * H, B, F, H, fk.op_mode);
                EMPOSE_CUDA_TRY(cudaGetLastError());
                EMPOSE_CUDA_TRY(cudaMemcpyAsync(lstm_state + (size_t)(LD + q) * B * H, pl.cstate[q], st_bytes, cudaMemcpyDeviceToDevice, s));
                ++ctx->launches;
            }
    EMPOSE_TRY(run(ctx, pl, pl.to_pose, mt_R, s));
    if (pose_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(pose_hat, pl.pose, (size_t)R * kPoseDim * 4, cudaMemcpyDeviceToDevice, s));
    if (!cfg.estimate_shape) return EMPOSE_OK;
    EMPOSE_TRY(run(ctx, pl, pl.to_shape, mt_R, s));
    // per-window shape mean (models.py:304-306) + pose features for the FK pass, through the LGD update kernel
    UpdateParams up;
    memset(&up, 0, sizeof(up));
    up.theta = pl.theta; up.beta = pl.beta; up.dtheta = pl.pose; up.dbeta = pl.dshape;
    up.step = 0.0f; up.first = 1; up.average_shape = cfg.average_shape; up.B = B; up.F = F; up.operand_mode = fk.op_mode;
    up.xiter = nullptr; up.pf = pl.pf; up.pf_stride = fk.pf_stride; up.pf_split = blend_operand_mode(&fk);
    EMPOSE_TRY(launch_update(up, s));
    ++ctx->launches;
    if (shape_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(shape_hat, pl.beta, (size_t)R * kBetas * 4, cudaMemcpyDeviceToDevice, s));
    if (!cfg.do_fk) return EMPOSE_OK;
    EMPOSE_TRY(run(ctx, pl, pl.pb, mt_R, s));
    identity_offsets_kernel<<<1, 160, 0, s>>>(pl.offsets);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    MainParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.sub = fk.sub; mp.fan = fk.fan; mp.spec = fk.spec; mp.theta = pl.theta; mp.vp = pl.vpoff; mp.jrest = pl.jrest;
    mp.offsets = pl.offsets; mp.rows_per_offset = R;       // one identity offset row for every frame
    mp.R = R; mp.want_grad = 0; mp.round_out = blend_operand_mode(&fk); mp.dj_ld = fk.dj_ld; mp.static_tree = fk.static_tree;
    mp.joints = pl.joints;
    EMPOSE_TRY(launch_main(mp, s));
    ctx->launches += 2;
    if (joints_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(joints_hat, pl.joints, (size_t)R * kPoseDim * 4, cudaMemcpyDeviceToDevice, s));
    return EMPOSE_OK;
}

}  // namespace
}  // namespace empose

extern "C" {
#pragma GCC visibility push(default)

int empose_rnn_create(const empose_rnn_config* cfg, const empose_tensor* tensors, int32_t n_tensors, empose_rnn** out) {
    if (!cfg || !tensors || !out) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    *out = nullptr;
    auto bad = [](const std::string& m) { set_last_error(m); return EMPOSE_E_ARG; };
    if (cfg->n_markers != 6 && cfg->n_markers != 12) return bad("n_markers must be 6 or 12 (reference models.py:116)");
    if (!cfg->use_marker_pos && !cfg->use_marker_ori) return bad("at least one of use_marker_pos / use_marker_ori is required");
    const int g = 4 * cfg->hidden_size;
    if (cfg->hidden_size < 8 || cfg->hidden_size % 8 || (g % 32) || (g > kMaxTileN && g % kMaxTileN))
        return bad("hidden_size must be a multiple of 8 that makes 4H a multiple of 32 and, above 256, of 256");
    if (cfg->num_layers < 1 || cfg->num_layers > 8) return bad("num_layers must be 1..8");
    if (cfg->learn_init_state) return bad("m_learn_init_state is not supported");
    if (cfg->do_fk && !cfg->estimate_shape) return bad("the FK pass needs a shape estimate (reference models.py:55)");
    if (cfg->estimate_shape && (cfg->shape_hidden_size < 16 || cfg->shape_hidden_size % 16)) return bad("shape_hidden_size must be a positive multiple of 16");
    if (cfg->precision < 0 || cfg->precision > 2) return bad("unknown precision");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device available: empose_b200 has no CPU fallback");
        return EMPOSE_E_CUDA;
    }
    EMPOSE_CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    EMPOSE_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) { set_last_error("empose_b200 is built for sm_100a (B200) only"); return EMPOSE_E_CUDA; }
    std::unique_ptr<empose_rnn> ctx(new empose_rnn());
    ctx->cfg = *cfg;
    IefData& fk = ctx->fk;
    fk.num_sms = prop.multiProcessorCount;
    fk.round = cfg->precision != EMPOSE_PRECISION_FP32;
    fk.op_mode = cfg->precision == EMPOSE_PRECISION_FP32 ? OPERAND_F32 : cfg->precision == EMPOSE_PRECISION_TF32 ? OPERAND_TF32 : OPERAND_F16;
    fk.op_half = fk.op_mode == OPERAND_F16 ? 1 : 0;
    configure_blend(&fk);
    fk.spec.use_pos = 0; fk.spec.use_ori = 0; fk.spec.weight = 1.0f;
    for (int i = 0; i < kSensors; ++i) fk.spec.sensor_active[i] = 0;
    ctx->dirs = cfg->bidirectional ? 2 : 1;
    ctx->n_pos = cfg->use_marker_pos ? 3 * cfg->n_markers : 0;
    ctx->in_size = ctx->n_pos + (cfg->use_marker_ori ? 9 * cfg->n_markers : 0);
    ctx->in_stride = round_up(ctx->in_size, fk.op_half ? 8 : 4);
    static const int kConfig6[6] = {0, 1, 2, 6, 7, 11};                     // reference configuration.py:89
    for (int i = 0; i < kSensors; ++i) ctx->slot_of_sensor[i] = cfg->n_markers == 12 ? i : -1;
    if (cfg->n_markers == 6) for (int i = 0; i < 6; ++i) ctx->slot_of_sensor[kConfig6[i]] = i;

    TensorTable tt{tensors, n_tensors};
    if (cfg->do_fk) EMPOSE_TRY(upload_submodel(&fk, tt));
    const int H = cfg->hidden_size, D = ctx->dirs;
    ctx->lstm.resize((size_t)cfg->num_layers * D);
    for (int l = 0; l < cfg->num_layers; ++l)
        for (int d = 0; d < D; ++d) {
            const int n_in = l == 0 ? ctx->in_size : D * H;
            const std::string sfx = "_l" + std::to_string(l) + (d == 1 ? "_reverse" : "");
            const float *wih, *whh, *bih, *bhh;
            EMPOSE_TRY(tt.get_f32("rnn.lstm.weight_ih" + sfx, {4 * H, n_in}, &wih));
            EMPOSE_TRY(tt.get_f32("rnn.lstm.weight_hh" + sfx, {4 * H, H}, &whh));
            EMPOSE_TRY(tt.get_f32("rnn.lstm.bias_ih" + sfx, {4 * H}, &bih));
            EMPOSE_TRY(tt.get_f32("rnn.lstm.bias_hh" + sfx, {4 * H}, &bhh));
            EMPOSE_TRY(pack_matrix(fk.arena, 4 * H, H, n_in, 32, fk.op_mode, true, [&](int n) {
                const int src = lstm_gate_of_packed(n) * H + lstm_unit_of_packed(n);
                return RowSource{whh + (size_t)src * H, wih + (size_t)src * n_in, 1.0, (double)bih[src] + (double)bhh[src]};
            }, &ctx->lstm[(size_t)l * D + d]));
        }
    if (fk.op_half) {
        ctx->wih_plain.resize((size_t)cfg->num_layers * D);
        ctx->whh_half.resize((size_t)cfg->num_layers * D);
        for (int l = 0; l < cfg->num_layers; ++l)
            for (int d = 0; d < D; ++d) {
                const int n_in = l == 0 ? ctx->in_size : D * H;
                const std::string sfx = "_l" + std::to_string(l) + (d == 1 ? "_reverse" : "");
                const float *wih, *whh, *bih, *bhh;
                EMPOSE_TRY(tt.get_f32("rnn.lstm.weight_ih" + sfx, {4 * H, n_in}, &wih));
                EMPOSE_TRY(tt.get_f32("rnn.lstm.weight_hh" + sfx, {4 * H, H}, &whh));
                EMPOSE_TRY(tt.get_f32("rnn.lstm.bias_ih" + sfx, {4 * H}, &bih));
                EMPOSE_TRY(tt.get_f32("rnn.lstm.bias_hh" + sfx, {4 * H}, &bhh));
                EMPOSE_TRY(pack_matrix(fk.arena, 4 * H, n_in, 0, 16, fk.op_mode, true, [&](int n) {
                    return RowSource{wih + (size_t)n * n_in, nullptr, 1.0, (double)bih[n] + (double)bhh[n]};
                }, &ctx->wih_plain[(size_t)l * D + d]));
                std::vector<__half> hh((size_t)4 * H * H);
                for (size_t i = 0; i < hh.size(); ++i) hh[i] = __float2half_rn(whh[i]);
                EMPOSE_TRY(fk.arena.upload(hh, &ctx->whh_half[(size_t)l * D + d]));
            }
    }
    EMPOSE_TRY(pack_linear(fk.arena, tt, "to_pose", "", "", kPoseDim, D * H, fk.op_mode, &ctx->to_pose));
    if (cfg->estimate_shape)
        EMPOSE_TRY(pack_mlp(fk.arena, tt, "to_shape", D * H, kBetas, cfg->shape_hidden_size, 2, false, fk.op_mode, &ctx->to_shape));
    *out = ctx.release();
    return EMPOSE_OK;
}

void empose_rnn_destroy(empose_rnn* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    delete ctx;
}

int empose_rnn_forward(empose_rnn* ctx, const float* marker_pos, const float* marker_oris, const int32_t* seq_lengths,
                       float* lstm_state, int32_t is_new_sequence, int32_t B, int32_t F, float* pose_hat, float* shape_hat,
                       float* joints_hat, void* stream) {
    if (!ctx) { set_last_error("null context"); return EMPOSE_E_ARG; }
    if (B < 1 || F < 1 || (int64_t)B * F > (int64_t)1 << 26) { set_last_error("B and F must be positive (and B*F <= 2^26)"); return EMPOSE_E_ARG; }
    if (!marker_pos || !marker_oris || !seq_lengths) { set_last_error("null input"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    RnnPlan* pl;
    EMPOSE_TRY(build_plan(ctx, B, F, &pl));
    return rnn_forward(ctx, *pl, marker_pos, marker_oris, seq_lengths, lstm_state, is_new_sequence, pose_hat, shape_hat, joints_hat,
                       static_cast<cudaStream_t>(stream));
}

int64_t empose_rnn_last_launch_count(const empose_rnn* ctx) { return ctx ? ctx->launches : 0; }

#pragma GCC visibility pop
}  // extern "C"
