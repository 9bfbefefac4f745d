// Model context, weight packing, execution plans and the C-ABI entry points (include/empose_b200.h).
//
// The pass implemented here is IterativeErrorFeedback.forward of the reference
// (empose/nn/models.py:485-632) in inference mode:
//
//   prepare  -> [LSTM (F+L-1 wavefront launches) -> heads | init MLPs] -> update(first)
//   N x { pose-blend GEMM -> frame kernel (SMPL sub-model fwd + reverse) -> transposed pose-blend GEMM
//         -> gradient features -> pose & shape iter-MLP chain -> update }
//   pose-blend GEMM -> frame kernel (forward only) -> outputs
//
// Every GEMM-shaped step is a list of GemmJob executed by the tcgen05 executor (TF32 mode) or the
// FFMA executor (FP32 mode).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/empose_b200.h"
#include "common.cuh"
#include "frame_kernels.h"
#include "gemm_jobs.h"
#include "gemm_tc.h"
#include "model_internal.h"

namespace empose {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

}  // namespace empose

using namespace empose;

namespace empose {
namespace {

// Evict least-recently-used plans, one at a time, until fewer than `limit` remain -- but never a plan the current C-ABI
// call has already used (last_use > call_clock): its kernels / copies may still be in flight.
template <typename Map>
void evict_lru(Map& plans, size_t limit, uint64_t call_clock) {
    while (plans.size() >= limit) {
        auto victim = plans.end();
        for (auto it = plans.begin(); it != plans.end(); ++it)
            if (it->second->last_use <= call_clock && (victim == plans.end() || it->second->last_use < victim->second->last_use)) victim = it;
        if (victim == plans.end()) return;          // everything belongs to this call: let the cache grow
        plans.erase(victim);
    }
}

int run_jobs(empose_ief* ctx, Plan& pl, const JobRange& r, int m_tiles, cudaStream_t s) {
    if (r.count == 0) return EMPOSE_OK;
    if (ctx->round) {
        ++ctx->last_launches;
        if (!ctx->profiling)
            return tc_launch(pl.book.d_jobs, pl.book.d_maps, r.begin, r.count, r.per_item, m_tiles, ctx->num_sms, s);
        if (ctx->prof_used == ctx->prof_events.size()) {
            cudaEvent_t a, b;
            EMPOSE_CUDA_TRY(cudaEventCreate(&a));
            EMPOSE_CUDA_TRY(cudaEventCreate(&b));
            ctx->prof_events.emplace_back(a, b);
        }
        auto& ev = ctx->prof_events[ctx->prof_used++];
        EMPOSE_CUDA_TRY(cudaEventRecord(ev.first, s));
        const int rc = tc_launch(pl.book.d_jobs, pl.book.d_maps, r.begin, r.count, r.per_item, m_tiles, ctx->num_sms, s);
        EMPOSE_CUDA_TRY(cudaEventRecord(ev.second, s));
        return rc;
    }
    return simt_launch(pl.book.d_jobs, pl.book.jobs.data(), r.begin, r.count, m_tiles, s, &ctx->last_launches);
}

// two interleaved MLP chains (pose, shape) reading the same input X
int build_mlp_pair(empose_ief* ctx, Plan& pl, const MlpPacked& mp, const MlpPacked& ms, const float* X, int64_t x_stride,
                   int x_k, JobRange* range) {
    const int R = pl.R, hidden = ctx->cfg.hidden_size;
    const int nl = (int)mp.layers.size();
    const MlpPacked* nets[2] = {&mp, &ms};
    int last_job[2] = {-1, -1};          // chain-local index of the last job of the previous layer
    float* finals[2] = {pl.dtheta, pl.dbeta};
    const int final_n[2] = {kPoseDim, kBetas};
    for (int l = 0; l < nl; ++l)
        for (int c = 0; c < 2; ++c) {
            const PackedMatrix& W = nets[c]->layers[l];
            ASrc a0, none;
            const bool scratch = ctx->round;
            const int hf = ctx->op_half;
            if (l == 0) a0 = ASrc{X, x_stride, x_k, R, hf};
            else a0 = ASrc{pl.act[c][(l - 1) & 1], hidden, hidden, pl.act_rows, hf};
            GemmJob proto;
            int m_rows = R;
            if (l == nl - 1) {
                proto = linear_proto(W, false, finals[c], final_n[c], final_n[c]);
            } else {
                proto = linear_proto(W, ctx->op_mode == OPERAND_TF32, pl.act[c][l & 1], hidden, hidden);
                proto.out_half = hf;
                if (ctx->cfg.skip_connections && l >= 2 && (l % 2) == 0) {   // end of a LinearLayers block
                    proto.res = pl.act[c][l & 1];                            // block input lives in the buffer being overwritten
                    proto.res_stride = hidden;
                }
                if (scratch) { proto.out_scratch = 1; m_rows = (int)pl.act_rows; }
            }
            if (scratch && l > 0) proto.a_scratch[0] = 1;
            EMPOSE_TRY(pl.book.add(W, a0, none, proto, m_rows, last_job[c], range));
            last_job[c] = range->count - 1;
        }
    range->per_item = range->count;
    return EMPOSE_OK;
}

int build_plan(empose_ief* ctx, int B, int F, Plan** out, int slot = 0) {
    auto key = std::make_tuple(B, F, slot);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) { it->second->last_use = ++ctx->plan_clock; *out = it->second.get(); return EMPOSE_OK; }
    evict_lru(ctx->plans, 8, ctx->call_clock);            // bound the workspace kept alive
    std::unique_ptr<Plan> plp(new Plan());
    Plan& pl = *plp;
    const empose_ief_config& cfg = ctx->cfg;
    pl.B = B; pl.F = F; pl.R = B * F;
    const int R = pl.R, hidden = cfg.hidden_size, H = cfg.rnn_hidden_size, L = cfg.rnn_num_layers, vp = ctx->sub.vp_dim;
    pl.book.use_tc = ctx->round;
    Arena& A = pl.arena;
    const size_t Rz = (size_t)R;
    EMPOSE_TRY(A.alloc_n(Rz * 144, &pl.meas));
    const int hf = ctx->op_half;
    const size_t esz = hf ? 2 : 4;                 // bytes per element of the MLP / LSTM operand buffers
    auto alloc_operand = [&](size_t elems, float** out) -> int {
        void* p;
        EMPOSE_TRY(A.alloc(elems * esz, &p, true));
        *out = static_cast<float*>(p);
        return EMPOSE_OK;
    };
    EMPOSE_TRY(alloc_operand(Rz * ctx->in_stride, &pl.xin));
    EMPOSE_TRY(alloc_operand(Rz * ctx->iter_stride, &pl.xiter));
    EMPOSE_TRY(A.alloc_n(Rz, &pl.coef));
    EMPOSE_TRY(A.alloc_n(Rz * kPoseDim, &pl.theta));
    EMPOSE_TRY(A.alloc_n(Rz * kBetas, &pl.beta));
    EMPOSE_TRY(A.alloc_n(Rz * kPoseDim, &pl.dtheta));
    EMPOSE_TRY(A.alloc_n(Rz * kBetas, &pl.dbeta));
    EMPOSE_TRY(A.alloc_n(Rz * ctx->pf_stride, &pl.pf, true));
    EMPOSE_TRY(A.alloc_n(Rz * vp, &pl.vpoff));
    EMPOSE_TRY(A.alloc_n(Rz * vp, &pl.dvp));
    EMPOSE_TRY(A.alloc_n(Rz * kPoseFeatPad, &pl.dpf));
    EMPOSE_TRY(A.alloc_n(Rz * kPoseDim, &pl.gth_part));
    EMPOSE_TRY(A.alloc_n(Rz * kJrestLd, &pl.jrest));
    EMPOSE_TRY(A.alloc_n(Rz * ctx->dj_ld, &pl.dj, true));
    EMPOSE_TRY(A.alloc_n((size_t)B * 144, &pl.offsets));
    EMPOSE_TRY(A.alloc_n(Rz * kPoseDim, &pl.joints));
    EMPOSE_TRY(A.alloc_n(Rz * 36, &pl.spos));
    EMPOSE_TRY(A.alloc_n(Rz * 108, &pl.sori));
    EMPOSE_TRY(A.alloc_n((size_t)B, &pl.seq_len));
    // Hidden activations of the MLP chains: in TF32 mode a CTA chains all layers of a 128-row tile, so the buffers
    // are CTA-local scratch (num_sms tiles, L2-resident); the FFMA executor runs layer by layer over all rows.
    pl.act_rows = ctx->round ? (int64_t)ctx->num_sms * kTileM : (int64_t)R;
    {
        const size_t one = ((size_t)pl.act_rows * hidden * esz + 1023) / 1024 * 1024;
        pl.act_block_bytes = 4 * one;
        EMPOSE_TRY(A.alloc(pl.act_block_bytes, &pl.act_block, true));
        for (int c = 0; c < 2; ++c)
            for (int q = 0; q < 2; ++q) pl.act[c][q] = reinterpret_cast<float*>(static_cast<char*>(pl.act_block) + (size_t)(2 * c + q) * one);
    }

    const int m_rows_R = R;
    if (cfg.rnn_init) {
        pl.hseq.resize(L); pl.cstate.resize(L); pl.hinit.resize(L);
        for (int l = 0; l < L; ++l) {
            EMPOSE_TRY(alloc_operand(Rz * H, &pl.hseq[l]));
            EMPOSE_TRY(A.alloc_n((size_t)B * H, &pl.cstate[l], true));
            EMPOSE_TRY(alloc_operand((size_t)B * H, &pl.hinit[l]));
        }
        pl.lstm_diag.resize(F + L - 1);
        const int mt_B = ceil_div(B, kTileM);
        if (ctx->round) EMPOSE_TRY(A.alloc_n((size_t)L * F * mt_B, &pl.wave_ctr, true));
        auto ctr_of = [&](int l, int t) { return pl.wave_ctr + ((size_t)l * F + t) * mt_B; };
        for (int d = 0; d < F + L - 1; ++d)
            for (int l = 0; l < L; ++l) {
                const int t = d - l;
                if (t < 0 || t >= F) continue;
                ASrc a0 = (t == 0) ? ASrc{pl.hinit[l], H, H, B, hf}
                                   : ASrc{operand_at(pl.hseq[l], (size_t)(t - 1) * H, hf), (int64_t)F * H, H, B, hf};
                ASrc a1 = (l == 0) ? ASrc{operand_at(pl.xin, (size_t)t * ctx->in_stride, hf), (int64_t)F * ctx->in_stride, ctx->in_size, B, hf}
                                   : ASrc{operand_at(pl.hseq[l - 1], (size_t)t * H, hf), (int64_t)F * H, H, B, hf};
                GemmJob proto;
                memset(&proto, 0, sizeof(proto));
                proto.epi = EPI_LSTM;
                proto.round_out = ctx->round ? 1 : 0;
                proto.out_half = hf;
                proto.out = operand_at(pl.hseq[l], (size_t)t * H, hf);
                proto.out_stride = (int64_t)F * H;
                proto.c_state = pl.cstate[l];
                proto.h_prev = a0.ptr;
                proto.h_prev_stride = a0.stride;
                proto.t = t;
                proto.hidden = H;
                proto.seq_len = pl.seq_len;
                proto.frames_per_window = 1;
                proto.split = 1 << 30;
                if (ctx->round) {
                    // ordering inside the persistent wavefront launch: step t of layer l reads h_{t-1} of its own layer (all
                    // N tiles of that step; its cell state and carried rows come with it) and h_t of the layer below
                    proto.done_ctr = ctr_of(l, t);
                    if (t > 0) { proto.wait_ctr[0] = ctr_of(l, t - 1); proto.wait_need[0] = (uint32_t)ctx->lstm[l].n_tiles; }
                    if (l > 0) { proto.wait_ctr[1] = ctr_of(l - 1, t); proto.wait_need[1] = (uint32_t)ctx->lstm[l - 1].n_tiles; }
                }
                EMPOSE_TRY(pl.book.add(ctx->lstm[l], a0, a1, proto, B, -1, &pl.lstm_diag[d]));
            }
        if (ctx->round) {
            // item table in wavefront order: diagonal by diagonal, row-tile units inside a diagonal, jobs inside a unit
            pl.wave_unit_rows = tc_item_rows(mt_B, ctx->num_sms);
            if (pl.wave_unit_rows < 0) return pl.wave_unit_rows;
            const int units = ceil_div(B, pl.wave_unit_rows);
            std::vector<int2> items;
            for (const JobRange& dg : pl.lstm_diag)
                for (int u = 0; u < units; ++u)
                    for (int j = dg.begin; j < dg.begin + dg.count; ++j) items.push_back(make_int2(j, u));
            int2* d_items;
            EMPOSE_TRY(A.upload(items, &d_items));
            pl.wave_items = d_items;
            pl.wave_n_items = (int)items.size();
        }
        // heads: theta0 | beta0 from the (masked) last-layer sequence
        GemmJob proto = linear_proto(ctx->heads, false, pl.dtheta, kPoseDim, kPoseDim + kBetas);
        proto.split = kPoseDim;
        proto.out2 = pl.dbeta;
        proto.out2_stride = kBetas;
        proto.mask_rows = R;                 // logical rows (the sequence lengths index windows of F of them)
        proto.seq_len = pl.seq_len;
        proto.frames_per_window = F;
        EMPOSE_TRY(pl.book.add(ctx->heads, ASrc{pl.hseq[L - 1], H, H, R, hf}, ASrc{}, proto, m_rows_R, -1, &pl.heads));
    } else {
        EMPOSE_TRY(build_mlp_pair(ctx, pl, ctx->pose_init, ctx->shape_init, pl.xin, ctx->in_stride, ctx->in_size, &pl.init_chain));
    }
    EMPOSE_TRY(build_mlp_pair(ctx, pl, ctx->pose_iter, ctx->shape_iter, pl.xiter, ctx->iter_stride, ctx->iter_in, &pl.iter_chain));
    EMPOSE_TRY(add_blend_jobs(pl.book, ctx, pl.pf, pl.vpoff, pl.jrest, R, &pl.pb));
    EMPOSE_TRY(add_blend_transposed_jobs(pl.book, ctx, pl.dvp, pl.dj, pl.dpf, R, &pl.pbt));
    EMPOSE_TRY(pl.book.finalize(A));
    *out = plp.get();
    plp->last_use = ++ctx->plan_clock;
    ctx->plans[key] = std::move(plp);
    return EMPOSE_OK;
}

int project_plan(empose_ief* ctx, int R, Plan** out) {
    auto it = ctx->project_plans.find(R);
    if (it != ctx->project_plans.end()) { it->second->last_use = ++ctx->plan_clock; *out = it->second.get(); return EMPOSE_OK; }
    evict_lru(ctx->project_plans, 4, ctx->call_clock);
    std::unique_ptr<Plan> plp(new Plan());
    Plan& pl = *plp;
    pl.R = R; pl.B = R; pl.F = 1;
    pl.book.use_tc = ctx->round;
    const int vp = ctx->sub.vp_dim;
    EMPOSE_TRY(pl.arena.alloc_n((size_t)R * ctx->pf_stride, &pl.pf, true));
    EMPOSE_TRY(pl.arena.alloc_n((size_t)R * vp, &pl.vpoff));
    EMPOSE_TRY(pl.arena.alloc_n((size_t)R * kJrestLd, &pl.jrest));
    EMPOSE_TRY(add_blend_jobs(pl.book, ctx, pl.pf, pl.vpoff, pl.jrest, R, &pl.pb));
    EMPOSE_TRY(pl.book.finalize(pl.arena));
    *out = plp.get();
    plp->last_use = ++ctx->plan_clock;
    ctx->project_plans[R] = std::move(plp);
    return EMPOSE_OK;
}

int check_config(const empose_ief_config& c) {
    auto bad = [](const std::string& m) { set_last_error(m); return EMPOSE_E_ARG; };
    if (c.n_markers != 6 && c.n_markers != 12) return bad("n_markers must be 6 or 12 (reference models.py:385)");
    if (!c.use_marker_pos && !c.use_marker_ori) return bad("at least one of use_marker_pos / use_marker_ori is required");
    if (c.num_iterations < 0 || c.num_iterations > 64) return bad("num_iterations out of range");
    if (c.hidden_size < 16 || c.hidden_size % 16) return bad("hidden_size must be a positive multiple of 16");
    if (c.num_layers < 0 || c.num_layers > 8) return bad("num_layers out of range");
    if (c.rnn_init) {
        const int g = 4 * c.rnn_hidden_size;
        if (c.rnn_hidden_size < 8 || (g % 32) || (g > kMaxTileN && g % kMaxTileN))
            return bad("rnn_hidden_size must make 4H a multiple of 32 and, above 256, of 256");
        if (c.rnn_num_layers < 1 || c.rnn_num_layers > 4) return bad("rnn_num_layers must be 1..4");
    }
    if (c.precision != EMPOSE_PRECISION_TF32 && c.precision != EMPOSE_PRECISION_FP32 && c.precision != EMPOSE_PRECISION_FP16)
        return bad("unknown precision");
    return EMPOSE_OK;
}

}  // namespace

int upload_submodel(IefData* ctx, const TensorTable& tt) {
    const int32_t* dims;
    EMPOSE_TRY(tt.get_i32("sub.dims", 6, &dims));   // n_verts, vp_dim, n_faces, max_degree, n_skin, n_sensors
    SubModel& m = ctx->sub;
    m.n_verts = dims[0]; m.vp_dim = dims[1]; m.n_faces = dims[2]; m.max_degree = dims[3]; m.n_skin = dims[4];
    if (dims[5] != kSensors) { set_last_error("exactly 12 sensors are supported"); return EMPOSE_E_ARG; }
    if (m.vp_dim > kMaxVp || m.vp_dim % 16 || m.n_verts * 3 > m.vp_dim || m.max_degree > kMaxDegree) {
        set_last_error("sub-model dimensions out of range");
        return EMPOSE_E_ARG;
    }
    auto up_f = [&](const char* name, std::initializer_list<int64_t> shape, const float** dst) -> int {
        const float* h;
        EMPOSE_TRY(tt.get_f32(name, shape, &h));
        int64_t n = 1;
        for (int64_t s : shape) n *= s;
        float* d;
        EMPOSE_TRY(ctx->arena.upload(std::vector<float>(h, h + n), &d));
        *dst = d;
        return EMPOSE_OK;
    };
    auto up_i = [&](const char* name, int64_t count, const int** dst, int lo, int hi) -> int {
        const int32_t* h;
        EMPOSE_TRY(tt.get_i32(name, count, &h));
        const empose_tensor* e = tt.find(name);
        const int64_t n = tt.numel(e);
        for (int64_t i = 0; i < n; ++i)
            if (h[i] < lo || h[i] >= hi) { set_last_error(std::string("index out of range in '") + name + "'"); return EMPOSE_E_ARG; }
        int* d;
        EMPOSE_TRY(ctx->arena.upload(std::vector<int>(h, h + n), &d));
        *dst = d;
        return EMPOSE_OK;
    };
    EMPOSE_TRY(up_f("sub.v_template", {m.vp_dim}, &m.v_template));
    EMPOSE_TRY(up_f("sub.shapedirs", {kBetas, m.vp_dim}, &m.shapedirs));
    EMPOSE_TRY(up_f("sub.j0", {kPoseDim}, &m.j0));
    EMPOSE_TRY(up_f("sub.jdirs", {kBetas, kPoseDim}, &m.jdirs));
    EMPOSE_TRY(up_f("sub.skin_weight", {m.n_verts, m.n_skin}, &m.skin_weight));
    EMPOSE_TRY(up_i("sub.skin_joint", (int64_t)m.n_verts * m.n_skin, &m.skin_joint, 0, kJoints));
    const int32_t* jt_ptr;
    EMPOSE_TRY(tt.get_i32("sub.jt_ptr", kJoints + 1, &jt_ptr));
    const int n_jt = jt_ptr[kJoints];
    EMPOSE_TRY(up_i("sub.jt_ptr", kJoints + 1, &m.jt_ptr, 0, n_jt + 1));
    EMPOSE_TRY(up_i("sub.jt_vert", n_jt, &m.jt_vert, 0, m.n_verts));
    EMPOSE_TRY(up_f("sub.jt_weight", {n_jt}, &m.jt_weight));
    const empose_tensor* vj = tt.find("sub.vj_ptr");
    if (!vj) { set_last_error("missing tensor 'sub.vj_ptr'"); return EMPOSE_E_MISSING; }
    m.n_vj = (int)tt.numel(vj) - 1;
    if (m.n_vj < 1 || m.n_vj > kMaxVj) { set_last_error("sub.vj_ptr: number of chunks out of range"); return EMPOSE_E_ARG; }
    EMPOSE_TRY(up_i("sub.vj_ptr", m.n_vj + 1, &m.vj_ptr, 0, n_jt + 1));
    EMPOSE_TRY(up_i("sub.jvj_ptr", kJoints + 1, &m.jvj_ptr, 0, m.n_vj + 1));
    const int32_t* vinc_ptr;
    EMPOSE_TRY(tt.get_i32("sub.vinc_ptr", m.n_verts + 1, &vinc_ptr));
    const int n_inc = vinc_ptr[m.n_verts];
    EMPOSE_TRY(up_i("sub.vinc_ptr", m.n_verts + 1, &m.vinc_ptr, 0, n_inc + 1));
    EMPOSE_TRY(up_i("sub.vinc_item", n_inc, &m.vinc_item, 0, kSensors * m.max_degree));
    EMPOSE_TRY(up_i("sub.vinc_code", n_inc, &m.vinc_code, 0, 5));
    EMPOSE_TRY(up_i("sub.parents", kJoints, &m.parents, -1, kJoints));
    {
        const int32_t* hp;
        EMPOSE_TRY(tt.get_i32("sub.parents", kJoints, &hp));
        ctx->static_tree = 1;
        for (int j = 0; j < kJoints; ++j)
            if (hp[j] != smpl_parent(j)) ctx->static_tree = 0;
    }
    EMPOSE_TRY(up_i("sub.faces", (int64_t)m.n_faces * 3, &m.faces, 0, m.n_verts));
    EMPOSE_TRY(up_i("sub.sensor_vert", kSensors, &m.sensor_vert, 0, m.n_verts));
    EMPOSE_TRY(up_i("sub.helper_vert", kSensors, &m.helper_vert, 0, m.n_verts));
    EMPOSE_TRY(up_i("sub.sensor_faces", (int64_t)kSensors * m.max_degree, &m.sensor_faces, -1, m.n_faces));
    EMPOSE_TRY(up_i("sub.sensor_degree", kSensors, &m.sensor_degree, 1, m.max_degree + 1));

    // ---- fan tables (csrc/fan_math.h; submodel.py) ----
    {
        FanModel& fm = ctx->fan;
        memset(&fm, 0, sizeof(fm));
        const int32_t* fd;
        EMPOSE_TRY(tt.get_i32("sub.fan_dims", 4, &fd));      // fan_ok, slots, max valence, partial sums per frame
        fm.ok = fd[0]; fm.slots = fd[1]; fm.max_deg = fd[2]; fm.n_part = fd[3];
        if (fm.ok && ((fm.slots != 8 && fm.slots != 12) || fm.max_deg >= fm.slots || fm.max_deg < 3 || fm.n_part < 1 ||
                      fm.n_part > kMaxPartials || kSensors * fm.slots != m.n_verts)) {
            set_last_error("sub.fan_dims out of range");
            return EMPOSE_E_ARG;
        }
        if (fm.ok) {
            fm.deg = m.sensor_degree;
            EMPOSE_TRY(up_i("sub.fan_helper", kSensors, &fm.helper, 0, fm.max_deg));
            EMPOSE_TRY(up_i("sub.fan_n_joints", kSensors, &fm.n_joints, 1, kMaxFanJoints + 1));
            EMPOSE_TRY(up_i("sub.fan_part_ptr", kSensors + 1, &fm.part_ptr, 0, fm.n_part + 1));
            EMPOSE_TRY(up_i("sub.fan_joint", (int64_t)kSensors * kMaxFanJoints, &fm.joint, 0, kJoints));
            EMPOSE_TRY(up_f("sub.fan_weight", {kSensors, kMaxFanJoints, fm.slots}, &fm.weight));
            const int32_t* jp;
            EMPOSE_TRY(tt.get_i32("sub.fan_jp_ptr", kJoints + 1, &jp));
            if (jp[kJoints] != fm.n_part) { set_last_error("sub.fan_jp_ptr does not cover the partial sums"); return EMPOSE_E_ARG; }
            EMPOSE_TRY(up_i("sub.fan_jp_ptr", kJoints + 1, &fm.jp_ptr, 0, fm.n_part + 1));
            EMPOSE_TRY(up_i("sub.fan_jp_idx", fm.n_part, &fm.jp_idx, 0, fm.n_part));
            const int32_t *hn, *hp;
            EMPOSE_TRY(tt.get_i32("sub.fan_n_joints", kSensors, &hn));
            EMPOSE_TRY(tt.get_i32("sub.fan_part_ptr", kSensors + 1, &hp));
            for (int i = 0; i < kSensors; ++i)
                if (hp[i + 1] - hp[i] != hn[i] || hp[0] != 0) { set_last_error("sub.fan_part_ptr is not the prefix sum of sub.fan_n_joints"); return EMPOSE_E_ARG; }
        }
    }

    // ---- the blend contraction and its transpose (kFeat* in frame_math.h) ----
    // forward: output column i < vp_dim is a blended-vertex coordinate, column vp_dim + c a rest-joint coordinate;
    //          K index k < 189 a pose feature, kFeatBeta + b a shape parameter.  W[i][k] = P[k][i] | S[b][i] | Jdirs[b][c];
    //          bias = v_template | J0 (added in fp32 after the accumulation).
    const float *P, *S, *VT, *J0, *JD;
    EMPOSE_TRY(tt.get_f32("sub.posedirs", {kPoseFeat, m.vp_dim}, &P));
    EMPOSE_TRY(tt.get_f32("sub.shapedirs", {kBetas, m.vp_dim}, &S));
    EMPOSE_TRY(tt.get_f32("sub.v_template", {m.vp_dim}, &VT));
    EMPOSE_TRY(tt.get_f32("sub.j0", {kPoseDim}, &J0));
    EMPOSE_TRY(tt.get_f32("sub.jdirs", {kBetas, kPoseDim}, &JD));
    const int vp = m.vp_dim, n_fwd = vp + kPoseDim;
    std::vector<float> wf((size_t)n_fwd * kFeatK, 0.0f), bf(n_fwd);
    for (int i = 0; i < n_fwd; ++i) {
        float* row = &wf[(size_t)i * kFeatK];
        if (i < vp) {
            for (int k = 0; k < kPoseFeat; ++k) row[k] = P[(size_t)k * vp + i];
            for (int b2 = 0; b2 < kBetas; ++b2) row[kFeatBeta + b2] = S[(size_t)b2 * vp + i];
            bf[i] = VT[i];
        } else {
            for (int b2 = 0; b2 < kBetas; ++b2) row[kFeatBeta + b2] = JD[(size_t)b2 * kPoseDim + (i - vp)];
            bf[i] = J0[i - vp];
        }
    }
    // power-of-two factor that brings the largest entry of a matrix to ~2^9 (fp16 operands: exact, no overflow, and the
    // 2^-11 of the split stays far above the subnormal range)
    auto pow2_scale = [](const float* w, size_t n) {
        float mx = 0.0f;
        for (size_t i = 0; i < n; ++i) mx = std::max(mx, std::fabs(w[i]));
        if (!(mx > 0.0f)) return 1.0f;
        return std::ldexp(1.0f, 9 - (int)std::ceil(std::log2(mx)));
    };
    auto f16r = [](float x) { return __half2float(__float2half_rn(x)); };
    if (ctx->blend_half) {
        // Error-compensated blend on fp16 operands (3xFP16), by K-concatenation like the tf32 form below: the feature row
        // holds x_hi = fp16(x) and x_lo = fp16((x - x_hi) 2^11); with W' = s W split the same way,
        //   [x_hi | x_lo | x_hi] . [W'_hi | W'_hi 2^-11 | W'_lo 2^-11]^T = s (x_hi W_hi + (x - x_hi) W_hi + x_hi (W - W_hi)),
        // accumulated in fp32 and multiplied by 1/s in the epilogue, where the fp32 bias (template / J0) is added.
        const float sc = pow2_scale(wf.data(), wf.size());
        std::vector<float> w0((size_t)n_fwd * 2 * kPoseFeatPad, 0.0f), w1((size_t)n_fwd * kFeatK, 0.0f);
        for (int i = 0; i < n_fwd; ++i)
            for (int k = 0; k < kFeatK; ++k) {
                const float v = wf[(size_t)i * kFeatK + k] * sc;
                const float hi = f16r(v);
                w0[(size_t)i * 2 * kPoseFeatPad + k] = hi;
                w0[(size_t)i * 2 * kPoseFeatPad + kPoseFeatPad + k] = f16r(hi / kSplitLoScale);
                w1[(size_t)i * kFeatK + k] = f16r(f16r((v - hi) * kSplitLoScale) / kSplitLoScale);
            }
        EMPOSE_TRY(pack_matrix(ctx->arena, n_fwd, 2 * kPoseFeatPad, kFeatK, 16, OPERAND_F16, true, [&](int r) {
            return RowSource{&w0[(size_t)r * 2 * kPoseFeatPad], &w1[(size_t)r * kFeatK], 1.0, (double)bf[r]};
        }, &ctx->pb));
        ctx->pb.out_scale = 1.0f / sc;
    } else if (ctx->round) {
        // Error-compensated (3xTF32) blend by K-concatenation: with x = x_hi + x_lo (both tf32),
        //   [x_hi | x_lo | x_hi] . [W_hi | W_hi | W_lo]^T = x_hi W_hi + x_lo W_hi + x_hi W_lo.
        // The vertex offsets feed cross products of ~1 cm mesh edges, which amplify plain TF32 rounding
        // into ~1e-3 rad of sensor-orientation noise; the split brings it back to fp32 level.
        std::vector<float> w0((size_t)n_fwd * 2 * kPoseFeatPad, 0.0f), w1((size_t)n_fwd * kFeatK, 0.0f);
        for (int i = 0; i < n_fwd; ++i)
            for (int k = 0; k < kFeatK; ++k) {
                const float v = wf[(size_t)i * kFeatK + k];
                const float hi = host_round_tf32(v);
                w0[(size_t)i * 2 * kPoseFeatPad + k] = hi;
                w0[(size_t)i * 2 * kPoseFeatPad + kPoseFeatPad + k] = hi;
                w1[(size_t)i * kFeatK + k] = host_round_tf32(v - hi);
            }
        EMPOSE_TRY(pack_matrix(ctx->arena, n_fwd, 2 * kPoseFeatPad, kFeatK, 16, OPERAND_TF32, true, [&](int r) {
            return RowSource{&w0[(size_t)r * 2 * kPoseFeatPad], &w1[(size_t)r * kFeatK], 1.0, (double)bf[r]};
        }, &ctx->pb));
    } else {
        EMPOSE_TRY(pack_matrix(ctx->arena, n_fwd, kFeatK, 0, 16, OPERAND_F32, true,
                               [&](int r) { return RowSource{&wf[(size_t)r * kFeatK], nullptr, 1.0, (double)bf[r]}; }, &ctx->pb));
    }
    // transposed: output column n < 189 is dE/dpf_n = sum_i P[n][i] dvp_i, column kFeatBeta + b is
    //             dE/dbeta_b = sum_i S[b][i] dvp_i + sum_c Jdirs[b][c] dJ_c; K = [dvp (vp_dim) | dJ (66)].
    // Single pass: tf32 operands, or (blend_half) fp16 operands -- the same 11 significant bits -- with the weights times a
    // power of two and dvp / dJ times kDvpScale, both undone by the epilogue.
    std::vector<float> zero_v(vp, 0.0f), zero_j(kPoseDim, 0.0f);
    float st = 1.0f;
    if (ctx->blend_half) st = std::min(std::min(pow2_scale(P, (size_t)kPoseFeat * vp), pow2_scale(S, (size_t)kBetas * vp)), pow2_scale(JD, (size_t)kBetas * kPoseDim));
    EMPOSE_TRY(pack_matrix(ctx->arena, kFeatK, vp, kPoseDim, 16, blend_operand_mode(ctx), false, [&](int r) {
        if (r < kPoseFeat) return RowSource{P + (size_t)r * vp, zero_j.data(), (double)st, 0.0};
        if (r < kFeatBeta) return RowSource{zero_v.data(), zero_j.data(), 1.0, 0.0};
        return RowSource{S + (size_t)(r - kFeatBeta) * vp, JD + (size_t)(r - kFeatBeta) * kPoseDim, (double)st, 0.0};
    }, &ctx->pbt));
    if (ctx->blend_half) ctx->pbt.out_scale = 1.0f / (st * kDvpScale);
    return EMPOSE_OK;
}

namespace {

int pack_lstm(empose_ief* ctx, const TensorTable& tt) {
    const int H = ctx->cfg.rnn_hidden_size, L = ctx->cfg.rnn_num_layers;
    ctx->lstm.resize(L);
    for (int l = 0; l < L; ++l) {
        const int n_in = l == 0 ? ctx->in_size : H;
        const std::string sfx = "_l" + std::to_string(l);
        const float *wih, *whh, *bih, *bhh;
        EMPOSE_TRY(tt.get_f32("rnn.lstm.weight_ih" + sfx, {4 * H, n_in}, &wih));
        EMPOSE_TRY(tt.get_f32("rnn.lstm.weight_hh" + sfx, {4 * H, H}, &whh));
        EMPOSE_TRY(tt.get_f32("rnn.lstm.bias_ih" + sfx, {4 * H}, &bih));
        EMPOSE_TRY(tt.get_f32("rnn.lstm.bias_hh" + sfx, {4 * H}, &bhh));
        // packed row n <- torch row gate*H + unit (gate order i,f,g,o; layers.py:114 / torch.nn.LSTM layout)
        EMPOSE_TRY(pack_matrix(ctx->arena, 4 * H, H, n_in, 32, ctx->op_mode, true, [&](int n) {
            const int src = lstm_gate_of_packed(n) * H + lstm_unit_of_packed(n);
            return RowSource{whh + (size_t)src * H, wih + (size_t)src * n_in, 1.0, (double)bih[src] + (double)bhh[src]};
        }, &ctx->lstm[l]));
    }
    // heads (models.py:429-430): rows 0..65 pose_net_init, 66..75 shape_net_init
    const float *wp, *bp, *ws, *bs;
    EMPOSE_TRY(tt.get_f32("pose_net_init.weight", {kPoseDim, H}, &wp));
    EMPOSE_TRY(tt.get_f32("pose_net_init.bias", {kPoseDim}, &bp));
    EMPOSE_TRY(tt.get_f32("shape_net_init.weight", {kBetas, H}, &ws));
    EMPOSE_TRY(tt.get_f32("shape_net_init.bias", {kBetas}, &bs));
    EMPOSE_TRY(pack_matrix(ctx->arena, kPoseDim + kBetas, H, 0, 16, ctx->op_mode, true, [&](int r) {
        if (r < kPoseDim) return RowSource{wp + (size_t)r * H, nullptr, 1.0, (double)bp[r]};
        return RowSource{ws + (size_t)(r - kPoseDim) * H, nullptr, 1.0, (double)bs[r - kPoseDim]};
    }, &ctx->heads));
    return EMPOSE_OK;
}

int forward_device(empose_ief* ctx, Plan& pl, const float* marker_pos, const float* marker_oris, const float* offset_r,
                   const float* offset_t, const int32_t* seq_lengths, const float* marker_masks, float* lstm_state,
                   int is_new_sequence, float* pose_hat, float* shape_hat, float* joints_hat,
                   const empose_ief_history* hist, cudaStream_t s) {
    const empose_ief_config& cfg = ctx->cfg;
    const int B = pl.B, F = pl.F, R = pl.R, H = cfg.rnn_hidden_size, L = cfg.rnn_num_layers, N = cfg.num_iterations;
    const int mt_R = ceil_div(R, kTileM), mt_B = ceil_div(B, kTileM);
    const int rnd = ctx->round ? 1 : 0;
    ctx->last_launches = 0;       // (the chunked host path adds the sub-batches up itself)
    auto count = [&](int rc) { ++ctx->last_launches; return rc; };

    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.seq_len, seq_lengths, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    EMPOSE_TRY(count(launch_pack_offsets(offset_r, offset_t, pl.offsets, B, s)));

    PrepareParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.marker_pos = marker_pos; pp.marker_oris = marker_oris; pp.seq_len = pl.seq_len; pp.masks = marker_masks;
    pp.R = R; pp.F = F;
    for (int i = 0; i < kSensors; ++i) pp.slot_of_sensor[i] = ctx->slot_of_sensor[i];
    pp.use_pos = cfg.use_marker_pos; pp.use_ori = cfg.use_marker_ori; pp.n_pos = ctx->n_pos;
    pp.in_size = ctx->in_size; pp.in_stride = ctx->in_stride; pp.iter_stride = ctx->iter_stride; pp.operand_mode = ctx->op_mode;
    pp.meas = pl.meas; pp.xin = pl.xin; pp.xiter = pl.xiter; pp.coef = pl.coef;
    EMPOSE_TRY(count(launch_prepare(pp, s)));

    if (cfg.rnn_init) {
        const size_t st_bytes = (size_t)B * H * 4;
        for (int l = 0; l < L; ++l) {
            if (lstm_state && !is_new_sequence) {
                EMPOSE_TRY(count(launch_to_operand(lstm_state + (size_t)l * B * H, pl.hinit[l], (int64_t)B * H, ctx->op_mode, s)));
                EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.cstate[l], lstm_state + (size_t)(L + l) * B * H, st_bytes, cudaMemcpyDeviceToDevice, s));
            } else {
                EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.hinit[l], 0, (size_t)B * H * operand_bytes(ctx->op_mode), s));
                EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.cstate[l], 0, st_bytes, s));
            }
        }
        if (ctx->round && pl.wave_items && debug_options().lstm_persistent) {
            // the whole wavefront (F + L - 1 diagonals) as ONE persistent launch; jobs order themselves through counters
            ++ctx->last_launches;
            cudaEvent_t ev0 = nullptr, ev1 = nullptr;
            if (ctx->profiling) {
                if (ctx->prof_used == ctx->prof_events.size()) {
                    cudaEvent_t a, b;
                    EMPOSE_CUDA_TRY(cudaEventCreate(&a));
                    EMPOSE_CUDA_TRY(cudaEventCreate(&b));
                    ctx->prof_events.emplace_back(a, b);
                }
                ev0 = ctx->prof_events[ctx->prof_used].first; ev1 = ctx->prof_events[ctx->prof_used].second;
                ++ctx->prof_used;
                EMPOSE_CUDA_TRY(cudaEventRecord(ev0, s));
            }
            EMPOSE_TRY(tc_launch_items(pl.book.d_jobs, pl.book.d_maps, pl.wave_items, pl.wave_n_items, pl.wave_unit_rows, ++pl.wave_epoch,
                                       mt_B, ctx->num_sms, s));
            if (ev1) EMPOSE_CUDA_TRY(cudaEventRecord(ev1, s));
        } else {
            for (const JobRange& d : pl.lstm_diag) EMPOSE_TRY(run_jobs(ctx, pl, d, mt_B, s));
        }
        EMPOSE_TRY(run_jobs(ctx, pl, pl.heads, mt_R, s));
        if (lstm_state)
            for (int l = 0; l < L; ++l) {
                EMPOSE_TRY(count(launch_gather_last(pl.hseq[l], lstm_state + (size_t)l * B * H, B, F, H, ctx->op_mode, s)));
                EMPOSE_CUDA_TRY(cudaMemcpyAsync(lstm_state + (size_t)(L + l) * B * H, pl.cstate[l], st_bytes, cudaMemcpyDeviceToDevice, s));
            }
    } else {
        EMPOSE_TRY(run_jobs(ctx, pl, pl.init_chain, mt_R, s));
    }

    for (int it = 0; it <= N; ++it) {
        UpdateParams up;
        memset(&up, 0, sizeof(up));
        up.theta = pl.theta; up.beta = pl.beta; up.dtheta = pl.dtheta; up.dbeta = pl.dbeta;
        up.step = cfg.step_size; up.first = (it == 0); up.average_shape = cfg.average_shape;
        up.B = B; up.F = F; up.operand_mode = ctx->op_mode;
        up.xiter = pl.xiter; up.in_size = ctx->in_size; up.iter_stride = ctx->iter_stride; up.pf = pl.pf;
        up.pf_stride = ctx->pf_stride; up.pf_split = blend_operand_mode(ctx);
        if (hist && hist->pose) up.hist_pose = hist->pose + (size_t)it * R * kPoseDim;
        if (hist && hist->shape) up.hist_shape = hist->shape + (size_t)it * R * kBetas;
        EMPOSE_TRY(count(launch_update(up, s)));
        const bool grad = (it < N) && cfg.use_gradient;
        // The final evaluation (models.py:593-609) only has to deliver the joints unless the caller keeps the marker
        // histories: the rest joints are the last columns of the blend GEMM (its last column tile), and the sub-model kernel
        // stops after the kinematic chain -- no ring skinning, no sensor frames.
        const bool joints_only = !grad && !(hist && (hist->markers || hist->markers_ori)) && ctx->fan.ok && !debug_options().main_general &&
                                 pl.pb.per_item == 1 && pl.pb.count > 1 && ctx->sub.vp_dim >= (pl.pb.count - 1) * kMaxTileN;
        if (joints_only) {
            JobRange last = pl.pb;
            last.begin += last.count - 1;
            last.count = 1;
            EMPOSE_TRY(run_jobs(ctx, pl, last, mt_R, s));
        } else {
            EMPOSE_TRY(run_jobs(ctx, pl, pl.pb, mt_R, s));
        }

        MainParams mp;
        memset(&mp, 0, sizeof(mp));
        mp.sub = ctx->sub; mp.fan = ctx->fan; mp.spec = ctx->spec;
        mp.theta = pl.theta; mp.vp = pl.vpoff; mp.jrest = pl.jrest;
        mp.offsets = pl.offsets; mp.rows_per_offset = F;
        mp.meas = pl.meas; mp.coef = pl.coef; mp.R = R; mp.want_grad = grad; mp.round_out = blend_operand_mode(ctx); mp.dj_ld = ctx->dj_ld;
        mp.static_tree = ctx->static_tree;
        mp.sensor_pos = (hist && hist->markers) ? hist->markers + (size_t)it * R * 36 : nullptr;
        mp.sensor_ori = (hist && hist->markers_ori) ? hist->markers_ori + (size_t)it * R * 108 : nullptr;
        mp.joints = (hist && hist->joints) ? hist->joints + (size_t)it * R * kPoseDim : nullptr;
        if (it == N && !mp.joints) mp.joints = pl.joints;
        mp.dvp = pl.dvp; mp.dj = pl.dj; mp.gtheta_part = pl.gth_part;
        static long long* tick_buf = nullptr;          // EMPOSE_MAIN_TICKS=1: dump phase timing of the first gradient launch
        static const bool want_ticks = getenv("EMPOSE_MAIN_TICKS") != nullptr;
        if (want_ticks && it == 0) {
            if (!tick_buf) cudaMalloc(&tick_buf, 32 * sizeof(long long));
            mp.ticks = tick_buf;
        }
        if (ctx->profiling) {
            if (ctx->prof_main_used == ctx->prof_main_events.size()) {
                cudaEvent_t a, b;
                EMPOSE_CUDA_TRY(cudaEventCreate(&a));
                EMPOSE_CUDA_TRY(cudaEventCreate(&b));
                ctx->prof_main_events.emplace_back(a, b);
            }
            auto& ev = ctx->prof_main_events[ctx->prof_main_used++];
            EMPOSE_CUDA_TRY(cudaEventRecord(ev.first, s));
            EMPOSE_TRY(count(launch_main(mp, s)));
            EMPOSE_CUDA_TRY(cudaEventRecord(ev.second, s));
        } else {
            EMPOSE_TRY(count(launch_main(mp, s)));
        }
        if (want_ticks && it == 0) {
            long long h[32];
            cudaMemcpy(h, tick_buf, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "main_kernel ticks:");
            for (int q = 1; q < 10; ++q) fprintf(stderr, " %lld", h[q] - h[q - 1]);
            fprintf(stderr, "\n");
        }
        if (it == N) {
            if (joints_hat)
                EMPOSE_CUDA_TRY(cudaMemcpyAsync(joints_hat, mp.joints, (size_t)R * kPoseDim * 4, cudaMemcpyDeviceToDevice, s));
            break;
        }
        if (grad) {
            EMPOSE_TRY(run_jobs(ctx, pl, pl.pbt, mt_R, s));
            PostParams po;
            memset(&po, 0, sizeof(po));
            po.theta = pl.theta; po.dpf = pl.dpf; po.gtheta_part = pl.gth_part; po.coef = pl.coef;
            po.R = R; po.operand_mode = ctx->op_mode; po.xiter = pl.xiter; po.in_size = ctx->in_size; po.iter_stride = ctx->iter_stride;
            EMPOSE_TRY(count(launch_post(po, s)));
        }
        // EMPOSE_L2_PERSIST=1 (experiment): the chained scratch activations as a persisting L2 access-policy window
        static const bool l2_persist = getenv("EMPOSE_L2_PERSIST") != nullptr;
        if (l2_persist && ctx->round) {
            static bool limit_set = false;
            if (!limit_set) {
                cudaDeviceProp prop;
                int dev = 0;
                cudaGetDevice(&dev);
                cudaGetDeviceProperties(&prop, dev);
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize);
                fprintf(stderr, "empose_b200: persisting L2 max %d MB, window max %d MB, scratch %zu MB\n", prop.persistingL2CacheMaxSize >> 20,
                        prop.accessPolicyMaxWindowSize >> 20, pl.act_block_bytes >> 20);
                limit_set = true;
            }
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = pl.act_block;
            attr.accessPolicyWindow.num_bytes = pl.act_block_bytes;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            EMPOSE_CUDA_TRY(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
        }
        EMPOSE_TRY(run_jobs(ctx, pl, pl.iter_chain, mt_R, s));
        if (l2_persist && ctx->round) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.num_bytes = 0;
            EMPOSE_CUDA_TRY(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
        }
    }
    if (pose_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(pose_hat, pl.theta, (size_t)R * kPoseDim * 4, cudaMemcpyDeviceToDevice, s));
    if (shape_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(shape_hat, pl.beta, (size_t)R * kBetas * 4, cudaMemcpyDeviceToDevice, s));
    return EMPOSE_OK;
}

int check_call(empose_ief* ctx, int B, int F) {
    if (!ctx) { set_last_error("null context"); return EMPOSE_E_ARG; }
    if (ctx->sensors_only && F != 1) { set_last_error("this context was made by empose_sensors_create: it only projects sensors"); return EMPOSE_E_ARG; }
    if (B < 1 || F < 1 || (int64_t)B * F > (int64_t)1 << 26) { set_last_error("B and F must be positive (and B*F <= 2^26)"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    ctx->call_clock = ctx->plan_clock;          // plans touched from here on belong to this call (evict_lru)
    return EMPOSE_OK;
}

}  // namespace
}  // namespace empose

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {
#pragma GCC visibility push(default)

int empose_abi_version(void) { return EMPOSE_ABI_VERSION; }
const char* empose_last_error(void) { return g_last_error.c_str(); }

int empose_set_option(const char* key, int32_t value) {
    if (!key) { set_last_error("null key"); return EMPOSE_E_ARG; }
    const std::string k(key);
    DebugOptions& o = debug_options();
    if (k == "main_general") o.main_general = value;
    else if (k == "fan_variant") o.fan_variant = value;
    else if (k == "lstm_persistent") o.lstm_persistent = value;
    else if (k == "blend_fp16") o.blend_fp16 = value;          // read when a context is created
    else { set_last_error("unknown option '" + k + "'"); return EMPOSE_E_ARG; }
    return EMPOSE_OK;
}

int empose_ief_create(const empose_ief_config* cfg, const empose_tensor* tensors, int32_t n_tensors, empose_ief** out) {
    if (!cfg || !tensors || !out) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    *out = nullptr;
    EMPOSE_TRY(check_config(*cfg));
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device available: empose_b200 has no CPU fallback");
        return EMPOSE_E_CUDA;
    }
    EMPOSE_CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    EMPOSE_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_last_error("empose_b200 is built for sm_100a (B200) only; found compute capability " + std::to_string(prop.major) +
                       "." + std::to_string(prop.minor));
        return EMPOSE_E_CUDA;
    }
    std::unique_ptr<empose_ief> ctx(new empose_ief());
    ctx->cfg = *cfg;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->round = cfg->precision != EMPOSE_PRECISION_FP32;
    ctx->op_mode = cfg->precision == EMPOSE_PRECISION_FP32 ? OPERAND_F32 : cfg->precision == EMPOSE_PRECISION_TF32 ? OPERAND_TF32 : OPERAND_F16;
    ctx->op_half = ctx->op_mode == OPERAND_F16 ? 1 : 0;
    configure_blend(ctx.get());
    ctx->n_pos = cfg->use_marker_pos ? 3 * cfg->n_markers : 0;
    ctx->in_size = ctx->n_pos + (cfg->use_marker_ori ? 9 * cfg->n_markers : 0);
    ctx->iter_in = ctx->in_size + kPoseDim + kBetas + (cfg->use_gradient ? kPoseDim + kBetas : 0);
    ctx->in_stride = round_up(ctx->in_size, ctx->op_half ? 8 : 4);        // rows must be 16-byte aligned for TMA
    ctx->iter_stride = round_up(ctx->iter_in, ctx->op_half ? 8 : 4);
    static const int kConfig6[6] = {0, 1, 2, 6, 7, 11};                     // reference configuration.py:89
    for (int i = 0; i < kSensors; ++i) { ctx->slot_of_sensor[i] = cfg->n_markers == 12 ? i : -1; }
    if (cfg->n_markers == 6) for (int i = 0; i < 6; ++i) ctx->slot_of_sensor[kConfig6[i]] = i;
    ctx->spec.use_pos = cfg->use_marker_pos; ctx->spec.use_ori = cfg->use_marker_ori; ctx->spec.weight = 1.0f;
    for (int i = 0; i < kSensors; ++i) ctx->spec.sensor_active[i] = ctx->slot_of_sensor[i] >= 0;

    TensorTable tt{tensors, n_tensors};
    EMPOSE_TRY(upload_submodel(ctx.get(), tt));
    const bool bn = cfg->batch_norm != 0;
    if (cfg->rnn_init) {
        EMPOSE_TRY(pack_lstm(ctx.get(), tt));
    } else {
        EMPOSE_TRY(pack_mlp(ctx->arena, tt, "pose_net_init", ctx->in_size, kPoseDim, cfg->hidden_size, cfg->num_layers, bn, ctx->op_mode, &ctx->pose_init));
        EMPOSE_TRY(pack_mlp(ctx->arena, tt, "shape_net_init", ctx->in_size, kBetas, cfg->hidden_size, cfg->num_layers, bn, ctx->op_mode, &ctx->shape_init));
    }
    EMPOSE_TRY(pack_mlp(ctx->arena, tt, "pose_net_iter", ctx->iter_in, kPoseDim, cfg->hidden_size, cfg->num_layers, bn, ctx->op_mode, &ctx->pose_iter));
    EMPOSE_TRY(pack_mlp(ctx->arena, tt, "shape_net_iter", ctx->iter_in, kBetas, cfg->hidden_size, cfg->num_layers, bn, ctx->op_mode, &ctx->shape_iter));
    *out = ctx.release();
    return EMPOSE_OK;
}

void empose_ief_destroy(empose_ief* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    delete ctx;
}

int64_t empose_ief_last_launch_count(const empose_ief* ctx) { return ctx ? ctx->last_launches : 0; }

int empose_ief_set_profiling(empose_ief* ctx, int32_t enable) {
    if (!ctx) { set_last_error("null context"); return EMPOSE_E_ARG; }
    ctx->profiling = enable != 0;
    ctx->prof_used = 0;
    ctx->prof_main_used = 0;
    return EMPOSE_OK;
}

int empose_ief_profile_read_main(empose_ief* ctx, double* main_ms, int64_t* main_launches) {
    if (!ctx || !main_ms || !main_launches) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    double total = 0.0;
    for (size_t i = 0; i < ctx->prof_main_used; ++i) {
        EMPOSE_CUDA_TRY(cudaEventSynchronize(ctx->prof_main_events[i].second));
        float ms = 0.0f;
        EMPOSE_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->prof_main_events[i].first, ctx->prof_main_events[i].second));
        total += ms;
    }
    *main_ms = total;
    *main_launches = (int64_t)ctx->prof_main_used;
    ctx->prof_main_used = 0;
    return EMPOSE_OK;
}

int empose_ief_profile_read(empose_ief* ctx, double* gemm_ms, int64_t* gemm_launches) {
    if (!ctx || !gemm_ms || !gemm_launches) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    double total = 0.0;
    for (size_t i = 0; i < ctx->prof_used; ++i) {
        EMPOSE_CUDA_TRY(cudaEventSynchronize(ctx->prof_events[i].second));
        float ms = 0.0f;
        EMPOSE_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->prof_events[i].first, ctx->prof_events[i].second));
        total += ms;
    }
    *gemm_ms = total;
    *gemm_launches = (int64_t)ctx->prof_used;
    ctx->prof_used = 0;
    return EMPOSE_OK;
}

int empose_ief_forward(empose_ief* ctx, const float* marker_pos, const float* marker_oris, const float* offset_r,
                       const float* offset_t, const int32_t* seq_lengths, const float* marker_masks, float* lstm_state,
                       int32_t is_new_sequence, int32_t B, int32_t F, float* pose_hat, float* shape_hat, float* joints_hat,
                       const empose_ief_history* history, void* stream) {
    EMPOSE_TRY(check_call(ctx, B, F));
    if (!marker_pos || !marker_oris || !offset_r || !offset_t || !seq_lengths) { set_last_error("null input"); return EMPOSE_E_ARG; }
    Plan* pl;
    EMPOSE_TRY(build_plan(ctx, B, F, &pl));
    return forward_device(ctx, *pl, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks, lstm_state,
                          is_new_sequence, pose_hat, shape_hat, joints_hat, history, static_cast<cudaStream_t>(stream));
}

// One sub-batch of the host-buffer entry point: windows [b0, b0 + bc) of a batch of B.  Uploads on `s_in`, runs the pass
// on `s`, downloads on `s_out`; the three streams are chained with events so that consecutive sub-batches overlap
// (upload k+1 | compute k | download k-1).
static int forward_host_chunk(empose_ief* ctx, Plan& pl, int B, int b0, const float* marker_pos, const float* marker_oris,
                              const float* offset_r, const float* offset_t, const int32_t* seq_lengths, const float* marker_masks,
                              float* lstm_state, int is_new_sequence, float* pose_hat, float* shape_hat, float* joints_hat,
                              const empose_ief_history* history, cudaStream_t s_in, cudaStream_t s, cudaStream_t s_out,
                              cudaEvent_t ev_in, cudaEvent_t ev_done) {
    const int bc = pl.B, F = pl.F;
    const size_t R = (size_t)pl.R, r0 = (size_t)b0 * F, Rall = (size_t)B * F;
    const int L = ctx->cfg.rnn_num_layers, H = ctx->cfg.rnn_hidden_size, N = ctx->cfg.num_iterations;
    const size_t state_n = ctx->cfg.rnn_init ? (size_t)2 * L * bc * H : 0;
    if (!pl.in_pos) {
        EMPOSE_TRY(pl.arena.alloc_n(R * 36, &pl.in_pos));
        EMPOSE_TRY(pl.arena.alloc_n(R * 108, &pl.in_ori));
        EMPOSE_TRY(pl.arena.alloc_n(R * 12, &pl.in_masks));
        EMPOSE_TRY(pl.arena.alloc_n(state_n, &pl.io_state));
        EMPOSE_TRY(pl.arena.alloc_n(R * kPoseDim, &pl.o_pose));
        EMPOSE_TRY(pl.arena.alloc_n(R * kBetas, &pl.o_shape));
        EMPOSE_TRY(pl.arena.alloc_n(R * kPoseDim, &pl.o_joints));
        EMPOSE_TRY(pl.arena.alloc_n((size_t)bc * 108, &pl.in_off_r));
        EMPOSE_TRY(pl.arena.alloc_n((size_t)bc * 36, &pl.in_off_t));
        EMPOSE_TRY(pl.arena.alloc_n((size_t)bc, &pl.in_len));
    }
    // ---- upload ----
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.in_pos, marker_pos + r0 * 36, R * 36 * 4, cudaMemcpyHostToDevice, s_in));
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.in_ori, marker_oris + r0 * 108, R * 108 * 4, cudaMemcpyHostToDevice, s_in));
    if (marker_masks) EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.in_masks, marker_masks + r0 * 12, R * 12 * 4, cudaMemcpyHostToDevice, s_in));
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.in_off_r, offset_r + (size_t)b0 * 108, (size_t)bc * 108 * 4, cudaMemcpyHostToDevice, s_in));
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.in_off_t, offset_t + (size_t)b0 * 36, (size_t)bc * 36 * 4, cudaMemcpyHostToDevice, s_in));
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.in_len, seq_lengths + b0, (size_t)bc * 4, cudaMemcpyHostToDevice, s_in));
    const bool with_state = lstm_state && state_n;
    if (with_state && !is_new_sequence)      // [2][L][B][H] on the host -> [2][L][bc][H] on the device
        for (int q = 0; q < 2 * L; ++q)
            EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.io_state + (size_t)q * bc * H, lstm_state + ((size_t)q * B + b0) * H, (size_t)bc * H * 4,
                                            cudaMemcpyHostToDevice, s_in));
    if (s_in != s) {
        EMPOSE_CUDA_TRY(cudaEventRecord(ev_in, s_in));
        EMPOSE_CUDA_TRY(cudaStreamWaitEvent(s, ev_in, 0));
    }
    // ---- compute ----
    empose_ief_history dh = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const size_t hist_dof[5] = {kPoseDim, kBetas, kPoseDim, 36, 108};
    float* hp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (history) {
        float* src[5] = {history->pose, history->shape, history->joints, history->markers, history->markers_ori};
        float** dp[5] = {&dh.pose, &dh.shape, &dh.joints, &dh.markers, &dh.markers_ori};
        for (int i = 0; i < 5; ++i) {
            hp[i] = src[i];
            if (hp[i]) {
                if (!pl.o_hist[i]) EMPOSE_TRY(pl.arena.alloc_n((size_t)(N + 1) * R * hist_dof[i], &pl.o_hist[i]));
                *dp[i] = pl.o_hist[i];
            }
        }
    }
    const int64_t launches_before = ctx->last_launches;
    EMPOSE_TRY(forward_device(ctx, pl, pl.in_pos, pl.in_ori, pl.in_off_r, pl.in_off_t, pl.in_len, marker_masks ? pl.in_masks : nullptr,
                              with_state ? pl.io_state : nullptr, is_new_sequence, pl.o_pose, pl.o_shape, pl.o_joints,
                              history ? &dh : nullptr, s));
    ctx->last_launches += launches_before;
    if (s_out != s) {
        EMPOSE_CUDA_TRY(cudaEventRecord(ev_done, s));
        EMPOSE_CUDA_TRY(cudaStreamWaitEvent(s_out, ev_done, 0));
    }
    // ---- download ----
    if (pose_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(pose_hat + r0 * kPoseDim, pl.o_pose, R * kPoseDim * 4, cudaMemcpyDeviceToHost, s_out));
    if (shape_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(shape_hat + r0 * kBetas, pl.o_shape, R * kBetas * 4, cudaMemcpyDeviceToHost, s_out));
    if (joints_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(joints_hat + r0 * kPoseDim, pl.o_joints, R * kPoseDim * 4, cudaMemcpyDeviceToHost, s_out));
    if (with_state)
        for (int q = 0; q < 2 * L; ++q)
            EMPOSE_CUDA_TRY(cudaMemcpyAsync(lstm_state + ((size_t)q * B + b0) * H, pl.io_state + (size_t)q * bc * H, (size_t)bc * H * 4,
                                            cudaMemcpyDeviceToHost, s_out));
    for (int i = 0; i < 5; ++i)
        if (hp[i])
            for (int it = 0; it <= N; ++it)      // [N+1][B][F][dof] on the host, [N+1][bc][F][dof] on the device
                EMPOSE_CUDA_TRY(cudaMemcpyAsync(hp[i] + ((size_t)it * Rall + r0) * hist_dof[i], pl.o_hist[i] + (size_t)it * R * hist_dof[i],
                                                R * hist_dof[i] * 4, cudaMemcpyDeviceToHost, s_out));
    return EMPOSE_OK;
}

int empose_ief_forward_host(empose_ief* ctx, const float* marker_pos, const float* marker_oris, const float* offset_r,
                            const float* offset_t, const int32_t* seq_lengths, const float* marker_masks, float* lstm_state,
                            int32_t is_new_sequence, int32_t B, int32_t F, float* pose_hat, float* shape_hat,
                            float* joints_hat, const empose_ief_history* history, void* stream) {
    EMPOSE_TRY(check_call(ctx, B, F));
    if (!marker_pos || !marker_oris || !offset_r || !offset_t || !seq_lengths) { set_last_error("null input"); return EMPOSE_E_ARG; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // Windows are independent, so a large batch is cut into sub-batches whose PCIe copies overlap the compute of their
    // neighbours.  Sub-batches stay >= 2048 windows: measured on the B200 at 4096 windows (profiles/r02/README.md), two
    // sub-batches take 11.08 ms against 11.54 ms unsplit (device time 8.68 ms, copies 2.8 ms), four take 12.1 ms -- below
    // 2048 windows the persistent LSTM wavefront (64 items per diagonal on 74 CTA pairs) and the last wave of the MLP chain
    // lose more than the overlap gains.
    static int min_chunk = 0;                       // EMPOSE_HOST_CHUNK=<windows> overrides the sub-batch size (experiments)
    if (min_chunk == 0) {
        const char* e = getenv("EMPOSE_HOST_CHUNK");
        min_chunk = e && atoi(e) > 0 ? atoi(e) : 2048;
    }
    const int n_chunks = std::max(1, std::min(4, B / min_chunk));
    ctx->last_launches = 0;
    if (n_chunks == 1) {
        Plan* plp;
        EMPOSE_TRY(build_plan(ctx, B, F, &plp));
        EMPOSE_TRY(forward_host_chunk(ctx, *plp, B, 0, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks, lstm_state,
                                      is_new_sequence, pose_hat, shape_hat, joints_hat, history, s, s, s, nullptr, nullptr));
        EMPOSE_CUDA_TRY(cudaStreamSynchronize(s));
        return EMPOSE_OK;
    }
    if (!ctx->copy_in) {
        EMPOSE_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        EMPOSE_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    }
    while ((int)ctx->pipe_events.size() < 2 * n_chunks + 2) {
        cudaEvent_t e;
        EMPOSE_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    // the copy streams start after whatever the caller already queued on `stream`
    cudaEvent_t ev_start = ctx->pipe_events[2 * n_chunks], ev_end = ctx->pipe_events[2 * n_chunks + 1];
    EMPOSE_CUDA_TRY(cudaEventRecord(ev_start, s));
    EMPOSE_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_in, ev_start, 0));
    EMPOSE_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_out, ev_start, 0));
    // Sub-batch sizes: the MLP chain works in waves of (CTA pairs) x 256 rows, so an even split of 4096 windows x 32 frames
    // (6.92 waves) into 2 x 3.46 costs a whole extra wave; cut at a multiple of one wave of windows instead (4 + 2.92).
    static const bool even_split = getenv("EMPOSE_HOST_EVEN") != nullptr;
    const int wave = std::max(1, (ctx->num_sms / 2) * 2 * kTileM / F);
    int b0 = 0;
    for (int c = 0; c < n_chunks; ++c) {
        int bc = B / n_chunks + (c < B % n_chunks ? 1 : 0);
        if (!even_split && wave < B / n_chunks) {
            if (c + 1 < n_chunks) bc = std::max(wave, (int)std::lround((double)(B / n_chunks) / wave) * wave);
            if (c + 1 == n_chunks || b0 + bc >= B) bc = B - b0;
        }
        if (bc <= 0) break;
        Plan* plp;
        EMPOSE_TRY(build_plan(ctx, bc, F, &plp, c + 1));        // one workspace per in-flight sub-batch
        EMPOSE_TRY(forward_host_chunk(ctx, *plp, B, b0, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks,
                                      lstm_state, is_new_sequence, pose_hat, shape_hat, joints_hat, history, ctx->copy_in, s,
                                      ctx->copy_out, ctx->pipe_events[2 * c], ctx->pipe_events[2 * c + 1]));
        b0 += bc;
    }
    EMPOSE_CUDA_TRY(cudaEventRecord(ev_end, ctx->copy_out));
    EMPOSE_CUDA_TRY(cudaStreamWaitEvent(s, ev_end, 0));
    EMPOSE_CUDA_TRY(cudaStreamSynchronize(s));
    return EMPOSE_OK;
}

int empose_ief_submit_host(empose_ief* ctx, const float* marker_pos, const float* marker_oris, const float* offset_r,
                           const float* offset_t, const int32_t* seq_lengths, const float* marker_masks, float* lstm_state,
                           int32_t is_new_sequence, int32_t B, int32_t F, float* pose_hat, float* shape_hat,
                           float* joints_hat, const empose_ief_history* history, int32_t slot, void* stream) {
    EMPOSE_TRY(check_call(ctx, B, F));
    if (!marker_pos || !marker_oris || !offset_r || !offset_t || !seq_lengths) { set_last_error("null input"); return EMPOSE_E_ARG; }
    if (slot < 0 || slot >= 4) { set_last_error("slot must be 0..3"); return EMPOSE_E_ARG; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!ctx->copy_in) {
        EMPOSE_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        EMPOSE_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    }
    for (auto& e : ctx->slot_events[slot])
        if (!e) EMPOSE_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // The upload goes on its own stream and is NOT ordered behind `stream`: it overlaps the compute of the requests before
    // it; the pass runs on `stream` once the upload is there; the download follows on a third stream.  Every slot has its
    // own workspace (plan), so up to four requests can be at different stages.
    Plan* plp;
    EMPOSE_TRY(build_plan(ctx, B, F, &plp, 8 + slot));
    ctx->last_launches = 0;
    EMPOSE_TRY(forward_host_chunk(ctx, *plp, B, 0, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks, lstm_state,
                                  is_new_sequence, pose_hat, shape_hat, joints_hat, history, ctx->copy_in, s, ctx->copy_out,
                                  ctx->slot_events[slot][0], ctx->slot_events[slot][1]));
    EMPOSE_CUDA_TRY(cudaEventRecord(ctx->slot_events[slot][2], ctx->copy_out));
    return EMPOSE_OK;
}

int empose_ief_wait_host(empose_ief* ctx, int32_t slot) {
    if (!ctx || slot < 0 || slot >= 4) { set_last_error("bad context or slot"); return EMPOSE_E_ARG; }
    if (!ctx->slot_events[slot][2]) { set_last_error("nothing was submitted into this slot"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaEventSynchronize(ctx->slot_events[slot][2]));
    return EMPOSE_OK;
}

int empose_sensors_create(const empose_tensor* tensors, int32_t n_tensors, int32_t precision, int32_t device, empose_ief** out) {
    if (!tensors || !out) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    *out = nullptr;
    if (precision < 0 || precision > 2) { set_last_error("unknown precision"); return EMPOSE_E_ARG; }
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
    std::unique_ptr<empose_ief> ctx(new empose_ief());
    memset(&ctx->cfg, 0, sizeof(ctx->cfg));
    ctx->cfg.device = device;
    ctx->cfg.precision = precision;
    ctx->sensors_only = true;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->round = precision != EMPOSE_PRECISION_FP32;
    ctx->op_mode = precision == EMPOSE_PRECISION_FP32 ? OPERAND_F32 : OPERAND_TF32;
    ctx->op_half = 0;
    configure_blend(ctx.get());
    ctx->spec.use_pos = 0; ctx->spec.use_ori = 0; ctx->spec.weight = 1.0f;
    for (int i = 0; i < kSensors; ++i) { ctx->spec.sensor_active[i] = 0; ctx->slot_of_sensor[i] = i; }
    TensorTable tt{tensors, n_tensors};
    EMPOSE_TRY(upload_submodel(ctx.get(), tt));
    *out = ctx.release();
    return EMPOSE_OK;
}

int empose_sensor_project(empose_ief* ctx, const float* poses, const float* shapes, const float* offset_r,
                          const float* offset_t, int32_t R, float* sensor_pos, float* sensor_ori, float* joints, void* stream) {
    EMPOSE_TRY(check_call(ctx, R, 1));
    if (!poses || !shapes || !offset_r || !offset_t) { set_last_error("null input"); return EMPOSE_E_ARG; }
    Plan* pl;
    EMPOSE_TRY(project_plan(ctx, R, &pl));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int rnd = ctx->round ? 1 : 0;
    ctx->last_launches = 1;
    EMPOSE_TRY(launch_pose_features(poses, shapes, pl->pf, ctx->pf_stride, blend_operand_mode(ctx), R, s));
    EMPOSE_TRY(run_jobs(ctx, *pl, pl->pb, ceil_div(R, kTileM), s));
    MainParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.sub = ctx->sub; mp.fan = ctx->fan; mp.spec = ctx->spec; mp.theta = poses; mp.vp = pl->vpoff; mp.jrest = pl->jrest;
    mp.offset_r = offset_r; mp.offset_t = offset_t; mp.rows_per_offset = 1; mp.R = R; mp.want_grad = 0; mp.round_out = blend_operand_mode(ctx);
    mp.dj_ld = ctx->dj_ld;
    mp.static_tree = ctx->static_tree;
    mp.sensor_pos = sensor_pos; mp.sensor_ori = sensor_ori; mp.joints = joints;
    ++ctx->last_launches;
    return launch_main(mp, s);
}

static int gemm_engine_run(int32_t precision, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                           float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t reps, float* ms_out,
                           cudaStream_t s) {
    if (!A || !W || !C || M < 1 || N < 1 || K < 1 || reps < 1) { set_last_error("bad argument"); return EMPOSE_E_ARG; }
    int dev = 0;
    EMPOSE_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    EMPOSE_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    PackedMatrix pm;
    choose_tiles(N, 16, &pm);
    Arena arena;
    JobBook book;
    book.use_tc = precision != EMPOSE_PRECISION_FP32;
    const int hf = precision == EMPOSE_PRECISION_FP16 ? 1 : 0;
    JobRange range;
    float* c_half = nullptr;                           // fp16 mode: operands converted here; C goes through an fp16 buffer
    const int64_t ldc_h = round_up(N, 8);              //            when N is a multiple of 32 (exercises the fp16 epilogue)
    const bool half_out = hf && (N % 32 == 0);
    if (hf) {
        const int64_t k8 = round_up(K, 8);
        float *a_h, *w_h;
        EMPOSE_TRY(arena.alloc((size_t)M * k8 * 2, reinterpret_cast<void**>(&a_h), true));
        EMPOSE_TRY(arena.alloc((size_t)N * k8 * 2, reinterpret_cast<void**>(&w_h), true));
        EMPOSE_TRY(launch_to_operand_2d(A, lda, M, K, a_h, k8, OPERAND_F16, s));
        EMPOSE_TRY(launch_to_operand_2d(W, ldw, N, K, w_h, k8, OPERAND_F16, s));
        A = a_h; lda = k8; W = w_h; ldw = k8;
        if (half_out) EMPOSE_TRY(arena.alloc((size_t)M * ldc_h * 2, reinterpret_cast<void**>(&c_half), true));
    }
    float* bias_padded = nullptr;                      // the epilogue reads bias in aligned groups of 32
    if (bias) {
        EMPOSE_TRY(arena.alloc_n((size_t)pm.n_pad + 32, &bias_padded, true));
        EMPOSE_CUDA_TRY(cudaMemcpyAsync(bias_padded, bias, (size_t)N * 4, cudaMemcpyDeviceToDevice, s));
    }
    GemmJob proto = half_out ? linear_proto(pm, false, c_half, ldc_h, N) : linear_proto(pm, false, C, ldc, N);
    proto.out_half = half_out ? 1 : 0;
    proto.in_half = hf;
    // the W tensor map describes the caller's matrix directly: K extent K (zero fill beyond), N rows
    for (int t = 0; t < pm.n_tiles; ++t) {
        GemmJob j = proto;
        j.a_ptr[0] = A; j.a_stride[0] = lda; j.a_k[0] = K;
        EMPOSE_TRY(book.get_map(A, lda, K, M, kTileM, hf, &j.a_map[0]));
        j.a_map[1] = -1;
        j.w_ptr = W; j.w_ld = ldw;
        EMPOSE_TRY(book.get_map(W, ldw, K, N, pm.tile_n, hf, &j.w_map));
        EMPOSE_TRY(book.get_map(W, ldw, K, N, pm.tile_n / 2, hf, &j.w_map2));
        j.n_begin = t * pm.tile_n; j.n_count = pm.tile_n; j.m_rows = M; j.dep = -1; j.bias = bias_padded;
        EMPOSE_TRY(book.attach_out_map(j));
        j.c_map1 = 0;
        if (range.count == 0) range.begin = (int)book.jobs.size();
        book.jobs.push_back(j);
        ++range.count;
    }
    EMPOSE_TRY(book.finalize(arena));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_out) {
        EMPOSE_CUDA_TRY(cudaEventCreate(&e0));
        EMPOSE_CUDA_TRY(cudaEventCreate(&e1));
    }
    int rc = EMPOSE_OK;
    for (int r = 0; r < reps + (ms_out ? 1 : 0) && rc == EMPOSE_OK; ++r) {
        if (ms_out && r == 1) cudaEventRecord(e0, s);           // launch 0 is the warm-up
        if (book.use_tc) rc = tc_launch(book.d_jobs, book.d_maps, range.begin, range.count, 1, ceil_div(M, kTileM), prop.multiProcessorCount, s);
        else rc = simt_launch(book.d_jobs, book.jobs.data(), range.begin, range.count, ceil_div(M, kTileM), s, nullptr);
    }
    if (ms_out) cudaEventRecord(e1, s);
    if (half_out && rc == EMPOSE_OK) rc = launch_from_operand_2d(c_half, ldc_h, M, N, C, ldc, OPERAND_F16, s);
    EMPOSE_CUDA_TRY(cudaStreamSynchronize(s));     // the job array is freed when `arena` goes out of scope
    if (ms_out) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out = ms / (float)reps;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return rc;
}

int empose_gemm_selftest(int32_t precision, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                         float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, void* stream) {
    return gemm_engine_run(precision, A, lda, W, ldw, bias, C, ldc, M, N, K, 1, nullptr, static_cast<cudaStream_t>(stream));
}

int empose_gemm_bench(int32_t precision, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                      float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t reps, float* ms_per_launch, void* stream) {
    if (!ms_per_launch) { set_last_error("null output"); return EMPOSE_E_ARG; }
    return gemm_engine_run(precision, A, lda, W, ldw, bias, C, ldc, M, N, K, reps, ms_per_launch, static_cast<cudaStream_t>(stream));
}

#pragma GCC visibility pop
}  // extern "C"
