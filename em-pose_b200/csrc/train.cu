// Training step of the LGD / IEF model (C ABI: empose_train_*, include/empose_b200.h).
//
// What the reference runs per step (scripts/train.py:136-149):
//   forward in train mode   IterativeErrorFeedback.forward  (empose/nn/models.py:485-632) with BatchNorm1d on
//                           batch statistics (layers.py:26,57) -- and, as a side effect, N backward passes of the
//                           reconstruction energy that already leave gradients in every upstream parameter (:576)
//   backward                IterativeErrorFeedback.backward (models.py:634-688): L1 pose / shape, reconstruction and
//                           FK losses over all N+1 iterates, total_loss.backward()
//
// Because the iter-MLP inputs are detached (models.py:549-551, 578-579) the only path from an iterate to the
// parameters is theta_i = theta_0 + step * sum_{k<i} dtheta_k, so the whole backward pass is
//   G_i   = d(everything that looks at iterate i)/d(theta_i, beta_i)      per frame (seed_kernel)
//   d(dtheta_k) = step * sum_{i>k} G_i,   d(theta_0) = sum_i G_i
// followed by plain MLP / heads / LSTM back-propagation.  Parameters live in ONE flat fp32 vector and their
// gradients in another (the layout is given by empose_train_layout), so data-parallel training needs exactly
// one all-reduce.  All GEMM-shaped work runs as GemmJob lists on the tcgen05 (TF32) or FFMA (FP32) executor:
//   forward    z = x W^T + b                       (W packed K-major from the flat vector every step)
//   backward   dx = dz W          via W^T packed   [in][out]
//              dW += dz^T x       via transposed copies dz^T [out][rows], x^T [in][rows] (contraction over rows)
#include <cuda_runtime.h>

#include <cmath>
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
#include "train_kernels.h"

namespace empose {
namespace {

constexpr float kBnEps = 1e-5f;        // torch.nn.BatchNorm1d defaults, as constructed at layers.py:26,57
constexpr float kBnMomentum = 0.1f;
constexpr int kInitBetaCol = 68;       // column of beta in the [theta | pad | beta] seed rows (16-byte aligned for TMA)
constexpr int kInitLd = 80;

struct Entry { std::string name; int kind; int64_t offset, numel; };   // kind 0: parameter, 1: BatchNorm running statistic

struct LinearSite { int n_in = 0, n_out = 0; int64_t w = -1, b = -1, gamma = -1, beta = -1, alpha = -1, rmean = -1, rvar = -1; };
struct MlpLayout { std::vector<LinearSite> layers; };
struct LstmLayout { std::vector<int64_t> wih, whh, bih, bhh; };

struct Layout {
    std::vector<Entry> entries;
    int64_t n_params = 0, n_buffers = 0;
    LstmLayout lstm;
    int64_t head_wp = -1, head_ws = -1, head_bp = -1, head_bs = -1;
    MlpLayout pose_init, shape_init, pose_iter, shape_iter;

    int64_t param(const std::string& name, int64_t numel) {
        const int64_t off = n_params;
        entries.push_back({name, 0, off, numel});
        n_params += (numel + 3) / 4 * 4;          // keep every tensor 16-byte aligned
        return off;
    }
    int64_t buffer(const std::string& name, int64_t numel) {
        const int64_t off = n_buffers;
        entries.push_back({name, 1, off, numel});
        n_buffers += (numel + 3) / 4 * 4;
        return off;
    }
    void mlp(const std::string& prefix, int n_in, int n_out, int hidden, int blocks, bool bn, MlpLayout* out) {
        auto site = [&](const std::string& lin, const std::string& bnn, const std::string& act, int in, int o) {
            LinearSite s;
            s.n_in = in; s.n_out = o;
            s.w = param(lin + ".weight", (int64_t)o * in);
            s.b = param(lin + ".bias", o);
            if (!bnn.empty()) {
                s.gamma = param(bnn + ".weight", o);
                s.beta = param(bnn + ".bias", o);
                s.rmean = buffer(bnn + ".running_mean", o);
                s.rvar = buffer(bnn + ".running_var", o);
            }
            if (!act.empty()) s.alpha = param(act + ".weight", 1);
            out->layers.push_back(s);
        };
        site(prefix + ".input_to_hidden", bn ? prefix + ".batch_norm" : "", prefix + ".activation_fn", n_in, hidden);
        const int stride = bn ? 4 : 3;
        for (int b = 0; b < blocks; ++b)
            for (int l = 0; l < 2; ++l) {
                const std::string base = prefix + ".hidden_layers." + std::to_string(b) + ".layers.";
                site(base + std::to_string(l * stride), bn ? base + std::to_string(l * stride + 1) : "",
                     base + std::to_string(l * stride + (bn ? 2 : 1)), hidden, hidden);
            }
        site(prefix + ".hidden_to_output", "", "", hidden, n_out);
    }
};

void make_layout(const empose_ief_config& c, Layout* L) {
    const int n_pos = c.use_marker_pos ? 3 * c.n_markers : 0;
    const int in_size = n_pos + (c.use_marker_ori ? 9 * c.n_markers : 0);
    const int iter_in = in_size + kPoseDim + kBetas + (c.use_gradient ? kPoseDim + kBetas : 0);
    const bool bn = c.batch_norm != 0;
    if (c.rnn_init) {
        const int H = c.rnn_hidden_size;
        for (int l = 0; l < c.rnn_num_layers; ++l) {
            const int n_in = l == 0 ? in_size : H;
            const std::string sfx = "_l" + std::to_string(l);
            L->lstm.wih.push_back(L->param("rnn.lstm.weight_ih" + sfx, (int64_t)4 * H * n_in));
            L->lstm.whh.push_back(L->param("rnn.lstm.weight_hh" + sfx, (int64_t)4 * H * H));
            L->lstm.bih.push_back(L->param("rnn.lstm.bias_ih" + sfx, 4 * H));
            L->lstm.bhh.push_back(L->param("rnn.lstm.bias_hh" + sfx, 4 * H));
        }
        L->head_wp = L->param("pose_net_init.weight", (int64_t)kPoseDim * H);
        L->head_bp = L->param("pose_net_init.bias", kPoseDim);
        L->head_ws = L->param("shape_net_init.weight", (int64_t)kBetas * H);
        L->head_bs = L->param("shape_net_init.bias", kBetas);
    } else {
        L->mlp("pose_net_init", in_size, kPoseDim, c.hidden_size, c.num_layers, bn, &L->pose_init);
        L->mlp("shape_net_init", in_size, kBetas, c.hidden_size, c.num_layers, bn, &L->shape_init);
    }
    L->mlp("pose_net_iter", iter_in, kPoseDim, c.hidden_size, c.num_layers, bn, &L->pose_iter);
    L->mlp("shape_net_iter", iter_in, kBetas, c.hidden_size, c.num_layers, bn, &L->shape_iter);
}

// a zero-initialised packed operand [n_pad][ld] (+ bias [n_pad + 32]) refreshed from the flat vector by PackOps
int alloc_packed(Arena& arena, int n, int k0, int k1, int granule, bool with_bias, PackedMatrix* pm) {
    choose_tiles(n, granule, pm);
    pm->kseg[0] = k0; pm->kseg[1] = k1;
    pm->koff[0] = 0; pm->koff[1] = round_up(k0, kChunkK);
    pm->ld = round_up(k0, kChunkK) + (k1 > 0 ? round_up(k1, kChunkK) : 0);
    EMPOSE_TRY(arena.alloc_n((size_t)pm->n_pad * pm->ld, &pm->w, true));
    if (with_bias) EMPOSE_TRY(arena.alloc_n((size_t)pm->n_pad + 32, &pm->bias, true));
    return EMPOSE_OK;
}

struct TrainMlp {
    const MlpLayout* lay = nullptr;
    std::vector<PackedMatrix> fw, bw;       // bw[l]: W_l^T [n_in][n_out] for l >= 1
};

// per-plan state of one MLP evaluated on S segments of R rows
struct MlpRun {
    const TrainMlp* net = nullptr;
    int S = 0;
    const float* X = nullptr; int64_t x_ld = 0; int x_k = 0;      // input rows [S*R][x_ld]
    float* out = nullptr; int64_t out_ld = 0;                      // final output [R][out_ld] (segment-independent buffer)
    float* D = nullptr; int64_t d_ld = 0;                          // seed dL/d(out) [S*R][d_ld]
    std::vector<float*> z, a, dzT, aT;                             // hidden layers 0 .. n_hidden-1
    std::vector<float*> mean, invstd;                              // [S][H] per hidden layer
    float *da = nullptr, *dz = nullptr, *DT = nullptr;
    float* dskip = nullptr;                                        // m_skip_connections: the gradient that bypasses a LinearLayers block
    const float* XT = nullptr;                                     // [x_k rows padded][ldT], shared by the two nets of a pair
    std::vector<std::vector<JobRange>> fwd;                        // [segment][layer]
    std::vector<JobRange> bwd_dx;                                  // [layer] (layer >= 1): da = dz_l W_l
};

// a weight-gradient GEMM, recorded while the plan is built and turned into jobs at the end (the jobs of one launch
// must be contiguous in the job array)
struct DwSpec { float* aT; int64_t ldT; int m_rows; float* wT; int n_in; int k; float* grad; int group; };

struct TrainPlan {
    int B = 0, F = 0, R = 0;
    std::vector<DwSpec> dw_specs;
    Arena arena;
    JobBook book;
    int64_t ldT = 0, ldT_iter = 0;           // row pitch of transposed [*][R] / [*][N*R] operands
    // forward workspace (same roles as Plan in model_internal.h)
    float *meas = nullptr, *xin = nullptr, *xiter = nullptr, *coef = nullptr;
    float *theta = nullptr, *beta = nullptr, *dtheta = nullptr, *dbeta = nullptr;
    float *pf = nullptr, *vpoff = nullptr, *dvp = nullptr, *dpf = nullptr, *gth_part = nullptr;
    float *jrest = nullptr, *dj = nullptr, *offsets = nullptr;
    cudaStream_t last_stream = nullptr;          // of the last backward pass (empose_train_loss_values)
    empose_loss_weights last_weights = {0.0f, 0.0f, 0.0f, 0.0f};
    bool last_fk = false, backward_done = false;
    int32_t* seq_len = nullptr;
    float* hist[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};     // pose, shape, joints, markers, markers_ori: [N+1][R][dof]
    float *g_theta = nullptr, *g_beta = nullptr;                         // [N+1][R][66|10]
    // LSTM
    std::vector<float*> hseq, cstate, hinit, gates, cseq, dgall, dx, dh_rec, dc_rec, dgT, hprevT, hT;
    float *xinT = nullptr, *dhead = nullptr;
    std::vector<JobRange> lstm_diag, bptt_diag;
    JobRange heads, heads_dx;
    // seeds
    float *d_dtheta = nullptr, *d_dbeta = nullptr, *d_init = nullptr, *d_init_masked = nullptr, *d_initT = nullptr;
    float *xiterT = nullptr;
    MlpRun pose_iter, shape_iter, pose_init, shape_init;
    JobRange pb, pbt, dw512, dw_small, dw_lstm;
    DwReduce* dw_red[3] = {nullptr, nullptr, nullptr};       // split-K reductions of the three weight-gradient launches
    int dw_red_n[3] = {0, 0, 0}, dw_red_max[3] = {0, 0, 0};
    double *stat_sums = nullptr, *col_scratch = nullptr, *loss_sums = nullptr;
    float* masks = nullptr;
    bool forward_done = false;
    bool have_masks = false;
};

}  // namespace
}  // namespace empose

using namespace empose;

struct empose_train {
    empose_ief* base = nullptr;        // sub-model, pose-blend operands, input geometry (weights inside are NOT used)
    empose_ief_config cfg;
    Layout layout;
    float *params = nullptr, *grads = nullptr, *bn_buffers = nullptr;
    empose_allreduce_fn sync_fn = nullptr;      // SyncBatchNorm: all-reduce of the batch statistics (null: per-rank statistics)
    void* sync_user = nullptr;
    int32_t sync_world = 1;
    Arena arena;
    std::vector<PackedMatrix> lstm_fw, lstm_bw;
    PackedMatrix heads_fw, heads_bw;
    TrainMlp pose_init, shape_init, pose_iter, shape_iter;
    std::vector<PackOp> pack_ops;
    PackOp* d_pack_ops = nullptr;
    int pack_max_elems = 0;
    std::unique_ptr<TrainPlan> plan;
    int64_t launches = 0;
    ~empose_train() {
        plan.reset();
        if (base) empose_ief_destroy(base);
    }
};

namespace empose {
namespace {

int run(empose_train* t, TrainPlan& pl, const JobRange& r, int m_tiles, cudaStream_t s) {
    if (r.count == 0) return EMPOSE_OK;
    if (t->base->round) {
        ++t->launches;
        return tc_launch(pl.book.d_jobs, pl.book.d_maps, r.begin, r.count, r.per_item, m_tiles, t->base->num_sms, s);
    }
    return simt_launch(pl.book.d_jobs, pl.book.jobs.data(), r.begin, r.count, m_tiles, s, &t->launches);
}

int reduce_dw(empose_train* t, TrainPlan& pl, int group, cudaStream_t s);
// SyncBatchNorm (empose_train_set_sync_batchnorm): all-reduce `count` doubles of statistics over the ranks, on the stream
int sync_sums(empose_train* t, double* sums, int64_t count, cudaStream_t s) {
    if (!t->sync_fn) return EMPOSE_OK;
    const int rc = t->sync_fn(t->sync_user, sums, count, s);
    if (rc != 0) { set_last_error("the SyncBatchNorm all-reduce callback failed"); return EMPOSE_E_ARG; }
    return EMPOSE_OK;
}
int64_t stat_rows(const empose_train* t, int R) { return t->sync_fn ? (int64_t)R * t->sync_world : (int64_t)R; }

// ---- packed operands and the PackOps that refresh them ---------------------------------------------------------
void add_op(empose_train* t, const float* src, const float* src2, float* dst, int rows, int cols, int64_t dst_ld, int64_t rs,
            int64_t cs, int lstm_map, bool round) {
    PackOp op;
    memset(&op, 0, sizeof(op));
    op.src = src; op.src2 = src2; op.dst = dst; op.rows = rows; op.cols = cols; op.dst_ld = dst_ld; op.src_rs = rs; op.src_cs = cs;
    op.lstm_map = lstm_map; op.hidden = t->cfg.rnn_hidden_size; op.round = round ? 1 : 0;
    t->pack_ops.push_back(op);
    if (rows * cols > t->pack_max_elems) t->pack_max_elems = rows * cols;
}

int build_mlp_operands(empose_train* t, const MlpLayout& lay, TrainMlp* net) {
    const bool rnd = t->base->round;
    net->lay = &lay;
    const int nl = (int)lay.layers.size();
    net->fw.resize(nl); net->bw.resize(nl);
    for (int l = 0; l < nl; ++l) {
        const LinearSite& s = lay.layers[l];
        EMPOSE_TRY(alloc_packed(t->arena, s.n_out, s.n_in, 0, 16, true, &net->fw[l]));
        add_op(t, t->params + s.w, nullptr, net->fw[l].w, s.n_out, s.n_in, net->fw[l].ld, s.n_in, 1, 0, rnd);
        add_op(t, t->params + s.b, nullptr, net->fw[l].bias, s.n_out, 1, 1, 1, 0, 0, false);
        if (l >= 1) {
            EMPOSE_TRY(alloc_packed(t->arena, s.n_in, s.n_out, 0, 16, false, &net->bw[l]));
            add_op(t, t->params + s.w, nullptr, net->bw[l].w, s.n_in, s.n_out, net->bw[l].ld, 1, s.n_in, 0, rnd);
        }
    }
    return EMPOSE_OK;
}

int build_operands(empose_train* t) {
    const empose_ief_config& c = t->cfg;
    const bool rnd = t->base->round;
    const Layout& L = t->layout;
    if (c.rnn_init) {
        const int H = c.rnn_hidden_size, nl = c.rnn_num_layers;
        t->lstm_fw.resize(nl); t->lstm_bw.resize(nl);
        for (int l = 0; l < nl; ++l) {
            const int n_in = l == 0 ? t->base->in_size : H;
            PackedMatrix& fw = t->lstm_fw[l];
            EMPOSE_TRY(alloc_packed(t->arena, 4 * H, H, n_in, 32, true, &fw));
            add_op(t, t->params + L.lstm.whh[l], nullptr, fw.w, 4 * H, H, fw.ld, H, 1, 1, rnd);
            add_op(t, t->params + L.lstm.wih[l], nullptr, fw.w + fw.koff[1], 4 * H, n_in, fw.ld, n_in, 1, 1, rnd);
            add_op(t, t->params + L.lstm.bih[l], t->params + L.lstm.bhh[l], fw.bias, 4 * H, 1, 1, 1, 0, 1, false);
            // backward operand: rows [h part (H) ; x part (n_in, layers >= 1)], K = 4H in torch gate order
            PackedMatrix& bw = t->lstm_bw[l];
            EMPOSE_TRY(alloc_packed(t->arena, l == 0 ? H : 2 * H, 4 * H, 0, 16, false, &bw));
            add_op(t, t->params + L.lstm.whh[l], nullptr, bw.w, H, 4 * H, bw.ld, 1, H, 0, rnd);
            if (l > 0) add_op(t, t->params + L.lstm.wih[l], nullptr, bw.w + (size_t)H * bw.ld, H, 4 * H, bw.ld, 1, n_in, 0, rnd);
        }
        PackedMatrix& hf = t->heads_fw;
        EMPOSE_TRY(alloc_packed(t->arena, kPoseDim + kBetas, H, 0, 16, true, &hf));
        add_op(t, t->params + L.head_wp, nullptr, hf.w, kPoseDim, H, hf.ld, H, 1, 0, rnd);
        add_op(t, t->params + L.head_ws, nullptr, hf.w + (size_t)kPoseDim * hf.ld, kBetas, H, hf.ld, H, 1, 0, rnd);
        add_op(t, t->params + L.head_bp, nullptr, hf.bias, kPoseDim, 1, 1, 1, 0, 0, false);
        add_op(t, t->params + L.head_bs, nullptr, hf.bias + kPoseDim, kBetas, 1, 1, 1, 0, 0, false);
        // dh = [d theta_0 | 0 0 | d beta_0] . Wh with K laid out like the seed rows
        PackedMatrix& hb = t->heads_bw;
        EMPOSE_TRY(alloc_packed(t->arena, H, kInitBetaCol + kBetas, 0, 16, false, &hb));
        add_op(t, t->params + L.head_wp, nullptr, hb.w, H, kPoseDim, hb.ld, 1, H, 0, rnd);
        add_op(t, t->params + L.head_ws, nullptr, hb.w + kInitBetaCol, H, kBetas, hb.ld, 1, H, 0, rnd);
    } else {
        EMPOSE_TRY(build_mlp_operands(t, L.pose_init, &t->pose_init));
        EMPOSE_TRY(build_mlp_operands(t, L.shape_init, &t->shape_init));
    }
    EMPOSE_TRY(build_mlp_operands(t, L.pose_iter, &t->pose_iter));
    EMPOSE_TRY(build_mlp_operands(t, L.shape_iter, &t->shape_iter));
    EMPOSE_TRY(t->arena.upload(t->pack_ops, &t->d_pack_ops));
    return EMPOSE_OK;
}

// ---- plan ------------------------------------------------------------------------------------------------------
// A "weight-like" operand living in an activation buffer: rows = output columns of the GEMM, K = contraction over rows
PackedMatrix operand_view(float* ptr, int n, int64_t ld, int k) {
    PackedMatrix pm;
    choose_tiles(n, 16, &pm);
    pm.w = ptr; pm.bias = nullptr; pm.ld = ld; pm.kseg[0] = k; pm.kseg[1] = 0; pm.koff[0] = 0; pm.koff[1] = 0;
    return pm;
}
int operand_rows(int n) { PackedMatrix pm; choose_tiles(n, 16, &pm); return pm.n_pad; }

GemmJob plain_proto(float* out, int64_t out_stride, int n_valid, bool round) {
    PackedMatrix dummy;
    return linear_proto(dummy, round, out, out_stride, n_valid);
}

// jobs accumulating dW[out][in] += A^T-operand . W-operand over K rows; group 0: 512-row outputs (hidden layers),
// 1: outputs of at most 128 rows (output layers, heads), 2: LSTM (4H rows)
void add_dw(TrainPlan& pl, float* aT, int64_t ldT, int m_rows, float* wT, int n_in, int k, float* grad, int group) {
    pl.dw_specs.push_back(DwSpec{aT, ldT, m_rows, wT, n_in, k, grad, group});
}
// Split-K: a weight gradient contracts over ALL rows (K = 65536 for the iter-MLPs at 512 windows x 32 frames x 4 iterations)
// into an output of a few tiles, so unsplit the 80 hidden-layer tiles ran two waves on 74 CTA pairs (the second one of 6) and
// the output-layer tiles ran on 8 SMs: 0.68 + 0.73 ms of a 15 ms step.  Every K slice writes its own partial product and one
// kernel adds the slices to the gradient in a fixed order.
constexpr int kDwSplits[3] = {4, 16, 1};
int emit_dw_jobs(TrainPlan& pl) {
    JobRange* ranges[3] = {&pl.dw512, &pl.dw_small, &pl.dw_lstm};
    for (int g = 0; g < 3; ++g) {
        std::vector<DwReduce> red;
        for (const DwSpec& d : pl.dw_specs) {
            if (d.group != g) continue;
            const int splits = pl.book.use_tc ? kDwSplits[g] : 1;
            if (splits == 1) {
                PackedMatrix W = operand_view(d.wT, d.n_in, d.ldT, d.k);
                GemmJob proto = plain_proto(d.grad, d.n_in, d.n_in, false);
                proto.res = d.grad;
                proto.res_stride = d.n_in;
                EMPOSE_TRY(pl.book.add(W, ASrc{d.aT, d.ldT, d.k, d.m_rows}, ASrc{}, proto, d.m_rows, -1, ranges[g]));
                continue;
            }
            const int ks = round_up(ceil_div(d.k, splits), 32);
            const int count = d.m_rows * d.n_in;
            float* partial;
            EMPOSE_TRY(pl.arena.alloc_n((size_t)splits * count, &partial, true));
            int used = 0;
            for (int i = 0; i < splits; ++i) {
                const int k0 = i * ks, kn = std::min(ks, d.k - k0);
                if (kn <= 0) break;
                PackedMatrix W = operand_view(d.wT, d.n_in, d.ldT, kn);
                W.koff[0] = k0;
                GemmJob proto = plain_proto(partial + (size_t)i * count, d.n_in, d.n_in, false);
                EMPOSE_TRY(pl.book.add(W, ASrc{d.aT + k0, d.ldT, kn, d.m_rows}, ASrc{}, proto, d.m_rows, -1, ranges[g]));
                ++used;
            }
            red.push_back(DwReduce{d.grad, partial, count, used});
            pl.dw_red_max[g] = std::max(pl.dw_red_max[g], count);
        }
        pl.dw_red_n[g] = (int)red.size();
        if (!red.empty()) EMPOSE_TRY(pl.arena.upload(red, &pl.dw_red[g]));
    }
    return EMPOSE_OK;
}

int reduce_dw(empose_train* t, TrainPlan& pl, int group, cudaStream_t s) {
    if (pl.dw_red_n[group] == 0) return EMPOSE_OK;
    ++t->launches;
    return launch_dw_reduce(pl.dw_red[group], pl.dw_red_n[group], pl.dw_red_max[group], s);
}

int build_mlp_run(empose_train* t, TrainPlan& pl, const TrainMlp& net, int S, const float* X, int64_t x_ld, int x_k,
                  const float* XT, float* out, int64_t out_ld, float* D, int64_t d_ld, MlpRun* run_out) {
    MlpRun& r = *run_out;
    const int R = pl.R, H = t->cfg.hidden_size;
    const bool rnd = t->base->round;
    const int nl = (int)net.lay->layers.size(), nh = nl - 1;
    const int64_t M = (int64_t)S * R, ldT = round_up((int)M, 32);
    Arena& A = pl.arena;
    r.net = &net; r.S = S; r.X = X; r.x_ld = x_ld; r.x_k = x_k; r.XT = XT; r.out = out; r.out_ld = out_ld; r.D = D; r.d_ld = d_ld;
    r.z.resize(nh); r.a.resize(nh); r.dzT.resize(nh); r.aT.resize(nh); r.mean.resize(nh); r.invstd.resize(nh);
    for (int l = 0; l < nh; ++l) {
        EMPOSE_TRY(A.alloc_n((size_t)M * H, &r.z[l]));
        EMPOSE_TRY(A.alloc_n((size_t)M * H, &r.a[l]));
        EMPOSE_TRY(A.alloc_n((size_t)operand_rows(H) * ldT, &r.dzT[l], true));
        EMPOSE_TRY(A.alloc_n((size_t)operand_rows(H) * ldT, &r.aT[l], true));
        EMPOSE_TRY(A.alloc_n((size_t)S * H, &r.mean[l], true));
        EMPOSE_TRY(A.alloc_n((size_t)S * H, &r.invstd[l], true));
    }
    EMPOSE_TRY(A.alloc_n((size_t)M * H, &r.da));
    EMPOSE_TRY(A.alloc_n((size_t)M * H, &r.dz));
    const bool skip = t->cfg.skip_connections != 0;
    if (skip) EMPOSE_TRY(A.alloc_n((size_t)M * H, &r.dskip));
    const int n_out = net.lay->layers[nl - 1].n_out;
    EMPOSE_TRY(A.alloc_n((size_t)kTileM * ldT, &r.DT, true));
    // forward jobs: one range per (segment, layer)
    r.fwd.assign(S, std::vector<JobRange>(nl));
    for (int k = 0; k < S; ++k)
        for (int l = 0; l < nl; ++l) {
            ASrc a0 = l == 0 ? ASrc{X + (size_t)k * R * x_ld, x_ld, x_k, R} : ASrc{r.a[l - 1] + (size_t)k * R * H, H, H, R};
            GemmJob proto = l == nl - 1 ? plain_proto(out, out_ld, n_out, false) : plain_proto(r.z[l] + (size_t)k * R * H, H, H, false);
            EMPOSE_TRY(pl.book.add(net.fw[l], a0, ASrc{}, proto, R, -1, &r.fwd[k][l]));
        }
    // backward dx jobs: da = (dz_l | D) . W_l for l = nl-1 .. 1
    r.bwd_dx.assign(nl, JobRange());
    for (int l = 1; l < nl; ++l) {
        ASrc a0 = l == nl - 1 ? ASrc{D, d_ld, n_out, M} : ASrc{r.dz, H, H, M};
        GemmJob proto = plain_proto(r.da, H, H, false);
        // LinearLayers skip (layers.py:35-43): blocks are the hidden layers (1, 2), (3, 4), ...; the gradient w.r.t. a block's
        // input is what comes back through its first layer PLUS what bypassed the block (saved in dskip at the block's end)
        if (skip && l >= 1 && l <= nl - 3 && (l % 2) == 1) { proto.res = r.dskip; proto.res_stride = H; }
        EMPOSE_TRY(pl.book.add(net.bw[l], a0, ASrc{}, proto, (int)M, -1, &r.bwd_dx[l]));
    }
    // dW jobs
    for (int l = 0; l < nl; ++l) {
        const LinearSite& s = net.lay->layers[l];
        float* wT = l == 0 ? const_cast<float*>(XT) : r.aT[l - 1];
        if (l == nl - 1) add_dw(pl, r.DT, ldT, n_out, wT, s.n_in, (int)M, t->grads + s.w, 1);
        else add_dw(pl, r.dzT[l], ldT, H, wT, s.n_in, (int)M, t->grads + s.w, 0);
    }
    (void)rnd;
    return EMPOSE_OK;
}

int build_plan(empose_train* t, int B, int F, TrainPlan** out) {
    if (t->plan && t->plan->B == B && t->plan->F == F) { *out = t->plan.get(); return EMPOSE_OK; }
    t->plan.reset();
    std::unique_ptr<TrainPlan> plp(new TrainPlan());
    TrainPlan& pl = *plp;
    const empose_ief* ctx = t->base;
    const empose_ief_config& cfg = t->cfg;
    pl.B = B; pl.F = F; pl.R = B * F;
    const int R = pl.R, H = cfg.rnn_hidden_size, L = cfg.rnn_num_layers, N = cfg.num_iterations, vp = ctx->sub.vp_dim;
    const bool rnd = ctx->round;
    pl.book.use_tc = rnd;
    Arena& A = pl.arena;
    const size_t Rz = (size_t)R;
    const int NS = N > 0 ? N : 1;
    pl.ldT = round_up(R, 32);
    pl.ldT_iter = round_up(NS * R, 32);
    EMPOSE_TRY(A.alloc_n(Rz * 144, &pl.meas));
    EMPOSE_TRY(A.alloc_n(Rz * ctx->in_stride, &pl.xin, true));
    EMPOSE_TRY(A.alloc_n(Rz * NS * ctx->iter_stride, &pl.xiter, true));
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
    EMPOSE_TRY(A.alloc_n((size_t)B, &pl.seq_len));
    EMPOSE_TRY(A.alloc_n(Rz * 12, &pl.masks));
    const size_t dof[5] = {kPoseDim, kBetas, kPoseDim, 36, 108};
    for (int i = 0; i < 5; ++i) EMPOSE_TRY(A.alloc_n((size_t)(N + 1) * Rz * dof[i], &pl.hist[i]));
    EMPOSE_TRY(A.alloc_n((size_t)(N + 1) * Rz * kPoseDim, &pl.g_theta, true));
    EMPOSE_TRY(A.alloc_n((size_t)(N + 1) * Rz * kBetas, &pl.g_beta, true));
    EMPOSE_TRY(A.alloc_n((size_t)NS * Rz * 68, &pl.d_dtheta, true));
    EMPOSE_TRY(A.alloc_n((size_t)NS * Rz * 12, &pl.d_dbeta, true));
    EMPOSE_TRY(A.alloc_n(Rz * kInitLd, &pl.d_init, true));
    EMPOSE_TRY(A.alloc_n(Rz * kInitLd, &pl.d_init_masked, true));
    EMPOSE_TRY(A.alloc_n((size_t)2 * 3 * std::max(cfg.hidden_size, 4 * H) * (NS + 1), &pl.stat_sums));
    EMPOSE_TRY(A.alloc_n((size_t)std::max(std::max(cfg.hidden_size, 4 * H), 128), &pl.col_scratch));
    EMPOSE_TRY(A.alloc_n((size_t)4, &pl.loss_sums));

    if (cfg.rnn_init) {
        auto rs = [&](std::vector<float*>& v) { v.assign(L, nullptr); };
        rs(pl.hseq); rs(pl.cstate); rs(pl.hinit); rs(pl.gates); rs(pl.cseq); rs(pl.dgall); rs(pl.dx); rs(pl.dh_rec); rs(pl.dc_rec);
        rs(pl.dgT); rs(pl.hprevT); rs(pl.hT);
        for (int l = 0; l < L; ++l) {
            EMPOSE_TRY(A.alloc_n(Rz * H, &pl.hseq[l], true));
            EMPOSE_TRY(A.alloc_n((size_t)B * H, &pl.cstate[l], true));
            EMPOSE_TRY(A.alloc_n((size_t)B * H, &pl.hinit[l], true));
            EMPOSE_TRY(A.alloc_n(Rz * 4 * H, &pl.gates[l], true));
            EMPOSE_TRY(A.alloc_n(Rz * H, &pl.cseq[l], true));
            EMPOSE_TRY(A.alloc_n(Rz * 4 * H, &pl.dgall[l], true));
            EMPOSE_TRY(A.alloc_n(Rz * H, &pl.dx[l], true));
            EMPOSE_TRY(A.alloc_n((size_t)B * H, &pl.dh_rec[l], true));
            EMPOSE_TRY(A.alloc_n((size_t)B * H, &pl.dc_rec[l], true));
            EMPOSE_TRY(A.alloc_n((size_t)operand_rows(4 * H) * pl.ldT, &pl.dgT[l], true));
            EMPOSE_TRY(A.alloc_n((size_t)operand_rows(H) * pl.ldT, &pl.hprevT[l], true));
            EMPOSE_TRY(A.alloc_n((size_t)operand_rows(H) * pl.ldT, &pl.hT[l], true));
        }
        EMPOSE_TRY(A.alloc_n((size_t)operand_rows(ctx->in_size) * pl.ldT, &pl.xinT, true));
        EMPOSE_TRY(A.alloc_n(Rz * H, &pl.dhead, true));
        EMPOSE_TRY(A.alloc_n((size_t)kTileM * pl.ldT, &pl.d_initT, true));
        // forward wavefront (as in model.cu) with the gates / cell sequence kept for the backward pass
        pl.lstm_diag.resize(F + L - 1);
        pl.bptt_diag.resize(F + L - 1);
        for (int d = 0; d < F + L - 1; ++d)
            for (int l = 0; l < L; ++l) {
                const int tt = d - l;
                if (tt < 0 || tt >= F) continue;
                ASrc a0 = (tt == 0) ? ASrc{pl.hinit[l], H, H, B} : ASrc{pl.hseq[l] + (size_t)(tt - 1) * H, (int64_t)F * H, H, B};
                ASrc a1 = (l == 0) ? ASrc{pl.xin + (size_t)tt * ctx->in_stride, (int64_t)F * ctx->in_stride, ctx->in_size, B}
                                   : ASrc{pl.hseq[l - 1] + (size_t)tt * H, (int64_t)F * H, H, B};
                GemmJob proto;
                memset(&proto, 0, sizeof(proto));
                proto.epi = EPI_LSTM;
                proto.round_out = rnd ? 1 : 0;
                proto.out = pl.hseq[l] + (size_t)tt * H;
                proto.out_stride = (int64_t)F * H;
                proto.c_state = pl.cstate[l];
                proto.h_prev = a0.ptr;
                proto.h_prev_stride = a0.stride;
                proto.t = tt;
                proto.hidden = H;
                proto.seq_len = pl.seq_len;
                proto.frames_per_window = 1;
                proto.split = 1 << 30;
                proto.gates_out = pl.gates[l] + (size_t)tt * 4 * H;
                proto.gates_stride = (int64_t)F * 4 * H;
                proto.c_seq_out = pl.cseq[l] + (size_t)tt * H;
                proto.c_seq_stride = (int64_t)F * H;
                EMPOSE_TRY(pl.book.add(t->lstm_fw[l], a0, a1, proto, B, -1, &pl.lstm_diag[d]));
            }
        // backward step of every cell: [dh_rec | dx] = dgates_t . [W_hh | W_ih]  (a separate loop: the jobs of one
        // launch must be contiguous in the job array)
        for (int d = 0; d < F + L - 1; ++d)
            for (int l = 0; l < L; ++l) {
                const int tt = d - l;
                if (tt < 0 || tt >= F) continue;
                if (tt == 0 && l == 0) continue;
                GemmJob bp = plain_proto(pl.dh_rec[l], H, l == 0 ? H : 2 * H, false);
                if (l > 0) { bp.split = H; bp.out2 = pl.dx[l - 1] + (size_t)tt * H; bp.out2_stride = (int64_t)F * H; }
                EMPOSE_TRY(pl.book.add(t->lstm_bw[l], ASrc{pl.dgall[l] + (size_t)tt * 4 * H, (int64_t)F * 4 * H, 4 * H, B}, ASrc{}, bp, B,
                                       -1, &pl.bptt_diag[d]));
            }
        GemmJob hp = plain_proto(pl.dtheta, kPoseDim, kPoseDim + kBetas, false);
        hp.split = kPoseDim; hp.out2 = pl.dbeta; hp.out2_stride = kBetas;
        hp.mask_rows = R; hp.seq_len = pl.seq_len; hp.frames_per_window = F;
        EMPOSE_TRY(pl.book.add(t->heads_fw, ASrc{pl.hseq[L - 1], H, H, R}, ASrc{}, hp, R, -1, &pl.heads));
        GemmJob hd = plain_proto(pl.dhead, H, H, false);
        EMPOSE_TRY(pl.book.add(t->heads_bw, ASrc{pl.d_init_masked, kInitLd, kInitBetaCol + kBetas, R}, ASrc{}, hd, R, -1, &pl.heads_dx));
        // dW of the heads and of the LSTM
        const Layout& LY = t->layout;
        add_dw(pl, pl.d_initT, pl.ldT, kPoseDim, pl.hT[L - 1], H, R, t->grads + LY.head_wp, 1);
        add_dw(pl, pl.d_initT + (size_t)kInitBetaCol * pl.ldT, pl.ldT, kBetas, pl.hT[L - 1], H, R, t->grads + LY.head_ws, 1);
        for (int l = 0; l < L; ++l) {
            add_dw(pl, pl.dgT[l], pl.ldT, 4 * H, pl.hprevT[l], H, R, t->grads + LY.lstm.whh[l], 2);
            if (l == 0) add_dw(pl, pl.dgT[l], pl.ldT, 4 * H, pl.xinT, ctx->in_size, R, t->grads + LY.lstm.wih[l], 2);
            else add_dw(pl, pl.dgT[l], pl.ldT, 4 * H, pl.hT[l - 1], H, R, t->grads + LY.lstm.wih[l], 2);
        }
    } else {
        EMPOSE_TRY(A.alloc_n((size_t)operand_rows(ctx->in_size) * pl.ldT, &pl.xinT, true));
        EMPOSE_TRY(build_mlp_run(t, pl, t->pose_init, 1, pl.xin, ctx->in_stride, ctx->in_size, pl.xinT, pl.dtheta, kPoseDim, pl.d_init,
                                 kInitLd, &pl.pose_init));
        EMPOSE_TRY(build_mlp_run(t, pl, t->shape_init, 1, pl.xin, ctx->in_stride, ctx->in_size, pl.xinT, pl.dbeta, kBetas,
                                 pl.d_init + kInitBetaCol, kInitLd, &pl.shape_init));
    }
    if (N > 0) {
        EMPOSE_TRY(A.alloc_n((size_t)operand_rows(ctx->iter_in) * pl.ldT_iter, &pl.xiterT, true));
        EMPOSE_TRY(build_mlp_run(t, pl, t->pose_iter, N, pl.xiter, ctx->iter_stride, ctx->iter_in, pl.xiterT, pl.dtheta, kPoseDim,
                                 pl.d_dtheta, 68, &pl.pose_iter));
        EMPOSE_TRY(build_mlp_run(t, pl, t->shape_iter, N, pl.xiter, ctx->iter_stride, ctx->iter_in, pl.xiterT, pl.dbeta, kBetas,
                                 pl.d_dbeta, 12, &pl.shape_iter));
    }
    EMPOSE_TRY(add_blend_jobs(pl.book, ctx, pl.pf, pl.vpoff, pl.jrest, R, &pl.pb));
    EMPOSE_TRY(add_blend_transposed_jobs(pl.book, ctx, pl.dvp, pl.dj, pl.dpf, R, &pl.pbt));
    EMPOSE_TRY(emit_dw_jobs(pl));
    EMPOSE_TRY(pl.book.finalize(A));
    *out = plp.get();
    t->plan = std::move(plp);
    return EMPOSE_OK;
}

// ---- MLP forward / backward on the job executors -------------------------------------------------------------
int mlp_forward_segment(empose_train* t, TrainPlan& pl, MlpRun* runs[2], int k, cudaStream_t s) {
    const int R = pl.R, H = t->cfg.hidden_size, mt = ceil_div(R, kTileM);
    const int rnd = t->base->round ? 1 : 0;
    const int nl = (int)runs[0]->net->lay->layers.size();
    for (int l = 0; l < nl; ++l) {
        for (int c = 0; c < 2; ++c) EMPOSE_TRY(run(t, pl, runs[c]->fwd[k][l], mt, s));
        if (l == nl - 1) break;
        for (int c = 0; c < 2; ++c) {
            MlpRun& r = *runs[c];
            const LinearSite& site = r.net->lay->layers[l];
            const float* z = r.z[l] + (size_t)k * R * H;
            float* mean = r.mean[l] + (size_t)k * H;
            float* istd = r.invstd[l] + (size_t)k * H;
            const bool bn = site.gamma >= 0;
            if (bn) {
                EMPOSE_TRY(launch_col_stats(z, H, R, 1, H, pl.stat_sums, s));
                EMPOSE_TRY(sync_sums(t, pl.stat_sums, 2 * (int64_t)H, s));
                EMPOSE_TRY(launch_bn_finalize(pl.stat_sums, stat_rows(t, R), 1, H, kBnEps, mean, istd, t->bn_buffers + site.rmean,
                                              t->bn_buffers + site.rvar, kBnMomentum, s));
                t->launches += 2;
            }
            EMPOSE_TRY(launch_bn_apply(z, H, R, 1, H, mean, istd, bn ? t->params + site.gamma : nullptr,
                                       bn ? t->params + site.beta : nullptr, t->params + site.alpha, rnd,
                                       r.a[l] + (size_t)k * R * H, H, s));
            ++t->launches;
            if (t->cfg.skip_connections && l >= 2 && (l % 2) == 0) {        // end of a LinearLayers block: + the block's input
                EMPOSE_TRY(launch_add_inplace(r.a[l] + (size_t)k * R * H, r.a[l - 2] + (size_t)k * R * H, (int64_t)R * H, rnd, s));
                ++t->launches;
            }
        }
    }
    return EMPOSE_OK;
}

int mlp_backward(empose_train* t, TrainPlan& pl, MlpRun& r, bool transpose_x, cudaStream_t s) {
    const int R = pl.R, H = t->cfg.hidden_size, S = r.S;
    const int64_t M = (int64_t)S * R, ldT = round_up((int)M, 32);
    const int mt = ceil_div((int)M, kTileM);
    const int rnd = t->base->round ? 1 : 0;
    const std::vector<LinearSite>& lay = r.net->lay->layers;
    const int nl = (int)lay.size(), n_out = lay[nl - 1].n_out;
    float* G = t->grads;
    // output layer: bias gradient from the exact seed, transposed (rounded) copy for dW, rounded copy for dx
    EMPOSE_TRY(launch_col_sum(r.D, r.d_ld, M, n_out, pl.col_scratch, G + lay[nl - 1].b, nullptr, s));
    EMPOSE_TRY(launch_transpose(r.D, r.d_ld, M, n_out, 0, 1, rnd, r.DT, ldT, s));
    t->launches += 3;
    if (rnd) { EMPOSE_TRY(launch_round_inplace(r.D, M, n_out, r.d_ld, s)); ++t->launches; }
    if (transpose_x) { EMPOSE_TRY(launch_transpose(r.X, r.x_ld, M, r.x_k, 0, 1, 0, const_cast<float*>(r.XT), ldT, s)); ++t->launches; }
    EMPOSE_TRY(run(t, pl, r.bwd_dx[nl - 1], mt, s));
    for (int l = nl - 2; l >= 0; --l) {
        const LinearSite& site = lay[l];
        const bool bn = site.gamma >= 0;
        const float* gamma = bn ? t->params + site.gamma : nullptr;
        const float* beta = bn ? t->params + site.beta : nullptr;
        const float* alpha = t->params + site.alpha;
        if (t->cfg.skip_connections && l >= 2 && (l % 2) == 0)             // da = dL/d(block output): it also reaches the block's input directly
            EMPOSE_CUDA_TRY(cudaMemcpyAsync(r.dskip, r.da, (size_t)M * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
        EMPOSE_TRY(launch_bn_bwd_reduce(r.da, H, r.z[l], H, R, S, H, r.mean[l], r.invstd[l], gamma, beta, alpha, pl.stat_sums, s));
        EMPOSE_TRY(launch_bn_param_grads(pl.stat_sums, S, H, bn ? G + site.gamma : nullptr, bn ? G + site.beta : nullptr, G + site.alpha, s));
        // SyncBatchNorm: the parameter gradients above come from the LOCAL sums (the flat gradient is averaged over the ranks
        // afterwards, as DDP does for torch.nn.SyncBatchNorm); dz needs the means of dy and dy * xhat over the GLOBAL batch
        if (bn) EMPOSE_TRY(sync_sums(t, pl.stat_sums, 3 * (int64_t)S * H, s));
        EMPOSE_TRY(launch_bn_bwd_apply(r.da, H, r.z[l], H, R, S, H, r.mean[l], r.invstd[l], gamma, beta, alpha, pl.stat_sums,
                                       bn ? stat_rows(t, R) : R, rnd, r.dz, H, s));
        // A bias in front of a BatchNorm has an analytically zero gradient (the batch mean removes it; the reference holds
        // ~1e-8 of rounding noise there): nothing is added.  Without BatchNorm it is the column sum of dz.
        if (!bn) EMPOSE_TRY(launch_col_sum(r.dz, H, M, H, pl.col_scratch, G + site.b, nullptr, s));
        EMPOSE_TRY(launch_transpose(r.dz, H, M, H, 0, 1, 0, r.dzT[l], ldT, s));
        EMPOSE_TRY(launch_transpose(r.a[l], H, M, H, 0, 1, 0, r.aT[l], ldT, s));     // operand of dW_{l+1}
        t->launches += 7;
        if (l >= 1) EMPOSE_TRY(run(t, pl, r.bwd_dx[l], mt, s));
    }
    return EMPOSE_OK;
}

void fill_main_params(const empose_train* t, const TrainPlan& pl, MainParams* mp) {
    const empose_ief* ctx = t->base;
    memset(mp, 0, sizeof(*mp));
    mp->sub = ctx->sub; mp->fan = ctx->fan; mp->spec = ctx->spec;
    mp->theta = pl.theta; mp->vp = pl.vpoff; mp->jrest = pl.jrest;
    mp->offsets = pl.offsets; mp->rows_per_offset = pl.F;
    mp->meas = pl.meas; mp->coef = pl.coef; mp->R = pl.R; mp->round_out = blend_operand_mode(ctx); mp->dj_ld = ctx->dj_ld;
    mp->static_tree = ctx->static_tree;
    mp->dvp = pl.dvp; mp->dj = pl.dj; mp->gtheta_part = pl.gth_part;
}

int gradient_tail(empose_train* t, TrainPlan& pl, int it, float* xiter, cudaStream_t s) {
    const empose_ief* ctx = t->base;
    EMPOSE_TRY(run(t, pl, pl.pbt, ceil_div(pl.R, kTileM), s));
    PostParams po;
    memset(&po, 0, sizeof(po));
    po.theta = pl.theta; po.dpf = pl.dpf; po.gtheta_part = pl.gth_part; po.coef = pl.coef;
    po.R = pl.R; po.operand_mode = ctx->op_mode; po.xiter = xiter; po.in_size = ctx->in_size; po.iter_stride = ctx->iter_stride;
    po.g_theta_out = pl.g_theta + (size_t)it * pl.R * kPoseDim;
    po.g_beta_out = pl.g_beta + (size_t)it * pl.R * kBetas;
    ++t->launches;
    return launch_post(po, s);
}

int train_forward(empose_train* t, TrainPlan& pl, const float* marker_pos, const float* marker_oris, const float* offset_r,
                  const float* offset_t, const int32_t* seq_lengths, const float* marker_masks, float* pose_hat, float* shape_hat,
                  float* joints_hat, const empose_ief_history* hist, cudaStream_t s) {
    const empose_ief* ctx = t->base;
    const empose_ief_config& cfg = t->cfg;
    const int B = pl.B, F = pl.F, R = pl.R, H = cfg.rnn_hidden_size, L = cfg.rnn_num_layers, N = cfg.num_iterations;
    const int mt_R = ceil_div(R, kTileM), mt_B = ceil_div(B, kTileM);
    const int rnd = ctx->round ? 1 : 0;
    t->launches = 0;
    pl.forward_done = false;
    EMPOSE_TRY(launch_pack(t->d_pack_ops, (int)t->pack_ops.size(), t->pack_max_elems, s));
    ++t->launches;
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.seq_len, seq_lengths, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    EMPOSE_TRY(launch_pack_offsets(offset_r, offset_t, pl.offsets, B, s));
    ++t->launches;
    pl.have_masks = marker_masks != nullptr;
    if (marker_masks) EMPOSE_CUDA_TRY(cudaMemcpyAsync(pl.masks, marker_masks, (size_t)R * 12 * 4, cudaMemcpyDeviceToDevice, s));

    for (int k = 0; k < (N > 0 ? N : 1); ++k) {        // every iteration keeps its own MLP input rows for the backward pass
        PrepareParams pp;
        memset(&pp, 0, sizeof(pp));
        pp.marker_pos = marker_pos; pp.marker_oris = marker_oris; pp.seq_len = pl.seq_len; pp.masks = marker_masks;
        pp.R = R; pp.F = F;
        for (int i = 0; i < kSensors; ++i) pp.slot_of_sensor[i] = ctx->slot_of_sensor[i];
        pp.use_pos = cfg.use_marker_pos; pp.use_ori = cfg.use_marker_ori; pp.n_pos = ctx->n_pos;
        pp.in_size = ctx->in_size; pp.in_stride = ctx->in_stride; pp.iter_stride = ctx->iter_stride; pp.operand_mode = ctx->op_mode;
        pp.meas = pl.meas; pp.xin = k == 0 ? pl.xin : nullptr; pp.xiter = pl.xiter + (size_t)k * R * ctx->iter_stride; pp.coef = pl.coef;
        EMPOSE_TRY(launch_prepare(pp, s));
        ++t->launches;
    }
    if (cfg.rnn_init) {
        for (int l = 0; l < L; ++l) {
            EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.hinit[l], 0, (size_t)B * H * 4, s));
            EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.cstate[l], 0, (size_t)B * H * 4, s));
        }
        for (const JobRange& d : pl.lstm_diag) EMPOSE_TRY(run(t, pl, d, mt_B, s));
        EMPOSE_TRY(run(t, pl, pl.heads, mt_R, s));
    } else {
        MlpRun* runs[2] = {&pl.pose_init, &pl.shape_init};
        EMPOSE_TRY(mlp_forward_segment(t, pl, runs, 0, s));
    }
    for (int it = 0; it <= N; ++it) {
        float* xiter_k = it < N ? pl.xiter + (size_t)it * R * ctx->iter_stride : nullptr;
        UpdateParams up;
        memset(&up, 0, sizeof(up));
        up.theta = pl.theta; up.beta = pl.beta; up.dtheta = pl.dtheta; up.dbeta = pl.dbeta;
        up.step = cfg.step_size; up.first = (it == 0); up.average_shape = cfg.average_shape;
        up.B = B; up.F = F; up.operand_mode = ctx->op_mode;
        up.xiter = xiter_k; up.in_size = ctx->in_size; up.iter_stride = ctx->iter_stride; up.pf = pl.pf;
        up.pf_stride = ctx->pf_stride; up.pf_split = blend_operand_mode(ctx);
        up.hist_pose = pl.hist[0] + (size_t)it * R * kPoseDim;
        up.hist_shape = pl.hist[1] + (size_t)it * R * kBetas;
        EMPOSE_TRY(launch_update(up, s));
        ++t->launches;
        EMPOSE_TRY(run(t, pl, pl.pb, mt_R, s));
        MainParams mp;
        fill_main_params(t, pl, &mp);
        mp.want_grad = it < N;                         // the final iterate's gradient needs the targets: see train_backward
        mp.joints = pl.hist[2] + (size_t)it * R * kPoseDim;
        mp.sensor_pos = pl.hist[3] + (size_t)it * R * 36;
        mp.sensor_ori = pl.hist[4] + (size_t)it * R * 108;
        EMPOSE_TRY(launch_main(mp, s));
        ++t->launches;
        if (it == N) break;
        EMPOSE_TRY(gradient_tail(t, pl, it, cfg.use_gradient ? xiter_k : nullptr, s));
        MlpRun* runs[2] = {&pl.pose_iter, &pl.shape_iter};
        EMPOSE_TRY(mlp_forward_segment(t, pl, runs, it, s));
    }
    const size_t dof[5] = {kPoseDim, kBetas, kPoseDim, 36, 108};
    if (pose_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(pose_hat, pl.theta, (size_t)R * kPoseDim * 4, cudaMemcpyDeviceToDevice, s));
    if (shape_hat) EMPOSE_CUDA_TRY(cudaMemcpyAsync(shape_hat, pl.beta, (size_t)R * kBetas * 4, cudaMemcpyDeviceToDevice, s));
    if (joints_hat)
        EMPOSE_CUDA_TRY(cudaMemcpyAsync(joints_hat, pl.hist[2] + (size_t)N * R * kPoseDim, (size_t)R * kPoseDim * 4, cudaMemcpyDeviceToDevice, s));
    if (hist) {
        float* hp[5] = {hist->pose, hist->shape, hist->joints, hist->markers, hist->markers_ori};
        for (int i = 0; i < 5; ++i)
            if (hp[i]) EMPOSE_CUDA_TRY(cudaMemcpyAsync(hp[i], pl.hist[i], (size_t)(N + 1) * R * dof[i] * 4, cudaMemcpyDeviceToDevice, s));
    }
    pl.forward_done = true;
    return EMPOSE_OK;
}

int read_losses(empose_train* t, TrainPlan& pl, float* loss_vals) {
    double sums[4];
    EMPOSE_CUDA_TRY(cudaMemcpyAsync(sums, pl.loss_sums, sizeof(sums), cudaMemcpyDeviceToHost, pl.last_stream));
    EMPOSE_CUDA_TRY(cudaStreamSynchronize(pl.last_stream));      // the reference synchronises here too (5 x .cpu().item(), models.py:676-680)
    const double n1 = (double)(t->cfg.num_iterations + 1);
    const empose_loss_weights& w = pl.last_weights;
    const double fk_sum = pl.last_fk ? sums[3] : 0.0;
    loss_vals[0] = (float)(sums[0] / n1);
    loss_vals[1] = (float)(sums[1] / n1);
    loss_vals[2] = (float)(sums[2] / n1);
    loss_vals[3] = (float)(fk_sum / n1);
    loss_vals[4] = (float)((w.pose_weight * sums[0] + w.fk_weight * fk_sum + w.shape_weight * sums[1] + w.reprojection_weight * sums[2]) / n1);
    return EMPOSE_OK;
}

int train_backward(empose_train* t, TrainPlan& pl, const float* poses_gt, const float* shapes_gt, const float* joints_gt,
                   const empose_loss_weights& w, float* loss_vals, cudaEvent_t dense_ready, cudaStream_t s) {
    const empose_ief* ctx = t->base;
    const empose_ief_config& cfg = t->cfg;
    const int B = pl.B, F = pl.F, R = pl.R, H = cfg.rnn_hidden_size, L = cfg.rnn_num_layers, N = cfg.num_iterations;
    const int mt_R = ceil_div(R, kTileM), mt_B = ceil_div(B, kTileM);
    const int rnd = ctx->round ? 1 : 0;
    const bool fk = w.fk_weight > 0.0f && joints_gt != nullptr;
    float* G = t->grads;
    bool dense_done = false;

    // ---- loss values (models.py:646-680) ----
    LossParams lp;
    memset(&lp, 0, sizeof(lp));
    lp.B = B; lp.F = F; lp.N = N; lp.seq_len = pl.seq_len; lp.coef = pl.coef;
    lp.pose_hist = pl.hist[0]; lp.shape_hist = pl.hist[1]; lp.markers_hist = pl.hist[3]; lp.markers_ori_hist = pl.hist[4];
    lp.joints_final = pl.hist[2] + (size_t)N * R * kPoseDim; lp.meas = pl.meas;
    lp.pose_gt = poses_gt; lp.shape_gt = shapes_gt; lp.joints_gt = fk ? joints_gt : nullptr;
    for (int i = 0; i < kSensors; ++i) lp.sensor_active[i] = ctx->spec.sensor_active[i];
    lp.use_pos = cfg.use_marker_pos; lp.use_ori = cfg.use_marker_ori; lp.sums = pl.loss_sums;
    EMPOSE_TRY(launch_losses(lp, s));
    ++t->launches;

    // ---- gradient of the final iterate: r_weight/(N+1) * reconstruction + fk_weight * FK (both times coef) ----
    {
        MainParams mp;
        fill_main_params(t, pl, &mp);
        mp.want_grad = 1;
        mp.spec.weight = w.reprojection_weight / (float)(N + 1);
        mp.joints_gt = fk ? joints_gt : nullptr;
        mp.joint_weight = w.fk_weight;
        EMPOSE_TRY(launch_main(mp, s));
        ++t->launches;
        EMPOSE_TRY(gradient_tail(t, pl, N, nullptr, s));
    }
    // ---- seeds ----
    SeedParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.B = B; sp.F = F; sp.N = N; sp.step = cfg.step_size; sp.average_shape = cfg.average_shape;
    sp.pose_w = w.pose_weight; sp.shape_w = w.shape_weight; sp.recon_w = w.reprojection_weight;
    sp.side_effect = cfg.use_gradient ? 1 : 0;
    sp.seq_len = pl.seq_len; sp.pose_hist = pl.hist[0]; sp.shape_hist = pl.hist[1]; sp.g_theta = pl.g_theta; sp.g_beta = pl.g_beta;
    sp.pose_gt = poses_gt; sp.shape_gt = shapes_gt;
    sp.d_dtheta = pl.d_dtheta; sp.ld_t = 68; sp.d_dbeta = pl.d_dbeta; sp.ld_b = 12;
    sp.d_init = pl.d_init; sp.ld_i = kInitLd; sp.init_beta_col = kInitBetaCol; sp.d_init_masked = pl.d_init_masked;
    EMPOSE_TRY(launch_seeds(sp, s));
    ++t->launches;

    // ---- iter MLPs ----
    if (N > 0) {
        EMPOSE_TRY(mlp_backward(t, pl, pl.pose_iter, true, s));
        EMPOSE_TRY(mlp_backward(t, pl, pl.shape_iter, false, s));
    }
    // ---- initial estimate ----
    if (cfg.rnn_init) {
        const Layout& LY = t->layout;
        EMPOSE_TRY(launch_col_sum(pl.d_init, kInitLd, R, kPoseDim, pl.col_scratch, G + LY.head_bp, nullptr, s));
        EMPOSE_TRY(launch_col_sum(pl.d_init + kInitBetaCol, kInitLd, R, kBetas, pl.col_scratch, G + LY.head_bs, nullptr, s));
        EMPOSE_TRY(launch_transpose(pl.d_init_masked, kInitLd, R, kInitBetaCol + kBetas, 0, 1, rnd, pl.d_initT, pl.ldT, s));
        t->launches += 5;
        if (rnd) { EMPOSE_TRY(launch_round_inplace(pl.d_init_masked, R, kInitBetaCol + kBetas, kInitLd, s)); ++t->launches; }
        EMPOSE_TRY(run(t, pl, pl.heads_dx, mt_R, s));
        // Everything the dense weight gradients (iter-MLPs, heads) contract is known now, and the transposes of the LSTM's
        // FORWARD data do not depend on the backward-through-time sweep: do both first, so that the dense bucket of the
        // flat gradient is final -- and can be all-reduced -- while the sweep runs.
        for (int l = 0; l < L; ++l) {
            EMPOSE_TRY(launch_transpose(pl.hseq[l], H, R, H, 1, F, 0, pl.hprevT[l], pl.ldT, s));
            EMPOSE_TRY(launch_transpose(pl.hseq[l], H, R, H, 0, 1, 0, pl.hT[l], pl.ldT, s));
            t->launches += 2;
        }
        EMPOSE_TRY(launch_transpose(pl.xin, ctx->in_stride, R, ctx->in_size, 0, 1, 0, pl.xinT, pl.ldT, s));
        ++t->launches;
        EMPOSE_TRY(run(t, pl, pl.dw512, ceil_div(cfg.hidden_size, kTileM), s));
        EMPOSE_TRY(run(t, pl, pl.dw_small, 1, s));
        EMPOSE_TRY(reduce_dw(t, pl, 0, s));
        EMPOSE_TRY(reduce_dw(t, pl, 1, s));
        dense_done = true;
        if (dense_ready) EMPOSE_CUDA_TRY(cudaEventRecord(dense_ready, s));
        for (int l = 0; l < L; ++l) {
            EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.dh_rec[l], 0, (size_t)B * H * 4, s));
            EMPOSE_CUDA_TRY(cudaMemsetAsync(pl.dc_rec[l], 0, (size_t)B * H * 4, s));
        }
        for (int d = F + L - 2; d >= 0; --d) {
            for (int l = 0; l < L; ++l) {
                const int tt = d - l;
                if (tt < 0 || tt >= F) continue;
                LstmCellBwdParams cp;
                memset(&cp, 0, sizeof(cp));
                cp.gates = pl.gates[l]; cp.c_seq = pl.cseq[l]; cp.dh_out = (l == L - 1) ? pl.dhead : pl.dx[l];
                cp.dh_rec = pl.dh_rec[l]; cp.dc_rec = pl.dc_rec[l]; cp.dgates = pl.dgall[l]; cp.seq_len = pl.seq_len;
                cp.B = B; cp.F = F; cp.H = H; cp.t = tt; cp.last = (tt == F - 1); cp.round_out = rnd;
                EMPOSE_TRY(launch_lstm_cell_bwd(cp, s));
                ++t->launches;
            }
            EMPOSE_TRY(run(t, pl, pl.bptt_diag[d], mt_B, s));
        }
        for (int l = 0; l < L; ++l) {
            EMPOSE_TRY(launch_col_sum(pl.dgall[l], 4 * H, R, 4 * H, pl.col_scratch, G + LY.lstm.bih[l], G + LY.lstm.bhh[l], s));
            EMPOSE_TRY(launch_transpose(pl.dgall[l], 4 * H, R, 4 * H, 0, 1, 0, pl.dgT[l], pl.ldT, s));
            t->launches += 3;
        }
    } else {
        EMPOSE_TRY(mlp_backward(t, pl, pl.pose_init, true, s));
        EMPOSE_TRY(mlp_backward(t, pl, pl.shape_init, false, s));
    }
    // ---- weight gradients: dW += dz^T x, contraction over the rows ----
    if (!dense_done) {
        EMPOSE_TRY(run(t, pl, pl.dw512, ceil_div(cfg.hidden_size, kTileM), s));
        EMPOSE_TRY(run(t, pl, pl.dw_small, 1, s));
        EMPOSE_TRY(reduce_dw(t, pl, 0, s));
        EMPOSE_TRY(reduce_dw(t, pl, 1, s));
        if (dense_ready) EMPOSE_CUDA_TRY(cudaEventRecord(dense_ready, s));
    }
    if (cfg.rnn_init) EMPOSE_TRY(run(t, pl, pl.dw_lstm, ceil_div(4 * H, kTileM), s));

    pl.last_stream = s; pl.last_weights = w; pl.last_fk = fk;
    if (loss_vals) EMPOSE_TRY(read_losses(t, pl, loss_vals));
    return EMPOSE_OK;
}

}  // namespace
}  // namespace empose

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {
#pragma GCC visibility push(default)

int empose_train_layout(const empose_ief_config* cfg, int32_t index, char* name_out, int32_t name_cap, int32_t* kind,
                        int64_t* offset, int64_t* numel) {
    if (!cfg || index < 0) { set_last_error("bad argument"); return EMPOSE_E_ARG; }
    Layout L;
    make_layout(*cfg, &L);
    if (index >= (int)L.entries.size()) return EMPOSE_E_MISSING;     // end of the list
    const Entry& e = L.entries[index];
    if (name_out && name_cap > 0) {
        strncpy(name_out, e.name.c_str(), (size_t)name_cap - 1);
        name_out[name_cap - 1] = '\0';
    }
    if (kind) *kind = e.kind;
    if (offset) *offset = e.offset;
    if (numel) *numel = e.numel;
    return EMPOSE_OK;
}

int empose_train_sizes(const empose_ief_config* cfg, int64_t* n_params, int64_t* n_buffers) {
    if (!cfg) { set_last_error("bad argument"); return EMPOSE_E_ARG; }
    Layout L;
    make_layout(*cfg, &L);
    if (n_params) *n_params = L.n_params;
    if (n_buffers) *n_buffers = L.n_buffers;
    return EMPOSE_OK;
}

int empose_train_create(const empose_ief_config* cfg, const empose_tensor* tensors, int32_t n_tensors, float* params,
                        float* grads, float* bn_buffers, empose_train** out) {
    if (!cfg || !tensors || !params || !grads || !out) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    *out = nullptr;
    if (cfg->precision == EMPOSE_PRECISION_FP16) {
        set_last_error("training runs in EMPOSE_PRECISION_TF32 or EMPOSE_PRECISION_FP32 (fp16 operands are an inference mode)");
        return EMPOSE_E_ARG;
    }
    std::unique_ptr<empose_train> t(new empose_train());
    t->cfg = *cfg;
    EMPOSE_TRY(empose_ief_create(cfg, tensors, n_tensors, &t->base));
    make_layout(*cfg, &t->layout);
    if (t->layout.n_buffers > 0 && !bn_buffers) { set_last_error("bn_buffers is required when BatchNorm is enabled"); return EMPOSE_E_ARG; }
    t->params = params; t->grads = grads; t->bn_buffers = bn_buffers;
    EMPOSE_TRY(build_operands(t.get()));
    *out = t.release();
    return EMPOSE_OK;
}

void empose_train_destroy(empose_train* t) {
    if (!t) return;
    cudaSetDevice(t->cfg.device);
    delete t;
}

int empose_train_forward(empose_train* t, const float* marker_pos, const float* marker_oris, const float* offset_r,
                         const float* offset_t, const int32_t* seq_lengths, const float* marker_masks, int32_t B, int32_t F,
                         float* pose_hat, float* shape_hat, float* joints_hat, const empose_ief_history* history, void* stream) {
    if (!t) { set_last_error("null context"); return EMPOSE_E_ARG; }
    if (B < 1 || F < 1 || (int64_t)B * F * (t->cfg.num_iterations > 0 ? t->cfg.num_iterations : 1) > ((int64_t)1 << 24)) {
        set_last_error("B and F must be positive and B*F*N <= 2^24 in training");
        return EMPOSE_E_ARG;
    }
    if (!marker_pos || !marker_oris || !offset_r || !offset_t || !seq_lengths) { set_last_error("null input"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(t->cfg.device));
    TrainPlan* pl;
    EMPOSE_TRY(build_plan(t, B, F, &pl));
    return train_forward(t, *pl, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks, pose_hat, shape_hat,
                         joints_hat, history, static_cast<cudaStream_t>(stream));
}

int empose_train_backward(empose_train* t, const float* poses_gt, const float* shapes_gt, const float* joints_gt,
                          const empose_loss_weights* weights, float* loss_vals, void* dense_ready_event, void* stream) {
    if (!t || !poses_gt || !shapes_gt || !weights) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    if (!t->plan || !t->plan->forward_done) { set_last_error("empose_train_backward needs a preceding empose_train_forward"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(t->cfg.device));
    const int rc = train_backward(t, *t->plan, poses_gt, shapes_gt, joints_gt, *weights, loss_vals,
                                  static_cast<cudaEvent_t>(dense_ready_event), static_cast<cudaStream_t>(stream));
    t->plan->forward_done = false;
    t->plan->backward_done = rc == EMPOSE_OK;
    return rc;
}

int empose_train_loss_values(empose_train* t, float* loss_vals) {
    if (!t || !loss_vals) { set_last_error("null argument"); return EMPOSE_E_ARG; }
    if (!t->plan || !t->plan->backward_done) { set_last_error("empose_train_loss_values needs a preceding empose_train_backward"); return EMPOSE_E_ARG; }
    EMPOSE_CUDA_TRY(cudaSetDevice(t->cfg.device));
    return read_losses(t, *t->plan, loss_vals);
}

int64_t empose_train_last_launch_count(const empose_train* t) { return t ? t->launches : 0; }

int empose_train_set_sync_batchnorm(empose_train* t, empose_allreduce_fn fn, void* user, int32_t world_size) {
    if (!t) { set_last_error("null context"); return EMPOSE_E_ARG; }
    if (fn && world_size < 1) { set_last_error("world_size must be >= 1"); return EMPOSE_E_ARG; }
    t->sync_fn = fn;
    t->sync_user = user;
    t->sync_world = fn ? world_size : 1;
    return EMPOSE_OK;
}

#pragma GCC visibility pop
}  // extern "C"
