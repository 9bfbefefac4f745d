// Host-side helpers shared by model.cu (the LGD model context) and smpl_full.cu (the full-mesh SMPL layer):
// tensor-table lookup, device arena, packed weight matrices, GEMM job book-keeping.
#pragma once

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
#include "gemm_jobs.h"
#include "frame_kernels.h"
#include "gemm_tc.h"

namespace empose {

#define EMPOSE_TRY(expr)            \
    do {                            \
        int _rc = (expr);           \
        if (_rc != EMPOSE_OK) return _rc; \
    } while (0)

inline float host_round_tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;   // inf / nan
    u += 0x1000u;                                       // round to nearest, ties away (cvt.rna)
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

// ---- tensor table -----------------------------------------------------------------------------
struct TensorTable {
    const empose_tensor* t;
    int n;
    const empose_tensor* find(const std::string& name) const {
        for (int i = 0; i < n; ++i)
            if (name == t[i].name) return &t[i];
        return nullptr;
    }
    int64_t numel(const empose_tensor* e) const {
        int64_t k = 1;
        for (int d = 0; d < e->ndim; ++d) k *= e->shape[d];
        return k;
    }
    // fetch a float tensor with an exact shape
    int get_f32(const std::string& name, std::initializer_list<int64_t> shape, const float** out) const {
        const empose_tensor* e = find(name);
        if (!e) { set_last_error("missing tensor '" + name + "'"); return EMPOSE_E_MISSING; }
        if (e->dtype != EMPOSE_F32) { set_last_error("tensor '" + name + "' must be float32"); return EMPOSE_E_SHAPE; }
        int64_t want = 1;
        for (int64_t s : shape) want *= s;
        if (numel(e) != want) {
            set_last_error("tensor '" + name + "' has " + std::to_string(numel(e)) + " elements, expected " + std::to_string(want));
            return EMPOSE_E_SHAPE;
        }
        *out = static_cast<const float*>(e->data);
        return EMPOSE_OK;
    }
    int get_i32(const std::string& name, int64_t count, const int32_t** out) const {
        const empose_tensor* e = find(name);
        if (!e) { set_last_error("missing tensor '" + name + "'"); return EMPOSE_E_MISSING; }
        if (e->dtype != EMPOSE_I32) { set_last_error("tensor '" + name + "' must be int32"); return EMPOSE_E_SHAPE; }
        if (count >= 0 && numel(e) != count) { set_last_error("tensor '" + name + "' has the wrong size"); return EMPOSE_E_SHAPE; }
        *out = static_cast<const int32_t*>(e->data);
        return EMPOSE_OK;
    }
};

// ---- device memory ------------------------------------------------------------------------------
struct Arena {
    std::vector<void*> ptrs;
    ~Arena() { for (void* p : ptrs) cudaFree(p); }
    int alloc(size_t bytes, void** out, bool zero = false) {
        void* p = nullptr;
        if (bytes == 0) bytes = 16;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            set_last_error("cudaMalloc of " + std::to_string(bytes) + " bytes failed");
            return EMPOSE_E_NOMEM;
        }
        ptrs.push_back(p);
        if (zero) EMPOSE_CUDA_TRY(cudaMemset(p, 0, bytes));
        *out = p;
        return EMPOSE_OK;
    }
    template <typename T> int alloc_n(size_t count, T** out, bool zero = false) {
        void* p;
        EMPOSE_TRY(alloc(count * sizeof(T), &p, zero));
        *out = static_cast<T*>(p);
        return EMPOSE_OK;
    }
    template <typename T> int upload(const std::vector<T>& h, T** out) {
        EMPOSE_TRY(alloc_n<T>(h.size(), out));
        if (!h.empty()) EMPOSE_CUDA_TRY(cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        return EMPOSE_OK;
    }
};

// ---- packed weights -----------------------------------------------------------------------------
struct PackedMatrix {
    int half = 0;             // 1: `w` holds fp16 elements (ld, koff, kseg count elements either way)
    float* w = nullptr;       // device [n_pad][ld]
    float* bias = nullptr;    // device [n_pad] or null
    int n = 0, n_pad = 0, tile_n = 0, n_tiles = 0;
    int kseg[2] = {0, 0};     // real K of each segment
    int koff[2] = {0, 0};     // column where the segment starts in w
    int64_t ld = 0;
    int has_act = 0;
    float alpha = 0.0f;
    float out_scale = 1.0f;   // GemmJob::out_scale of the jobs made from this matrix
};

inline void choose_tiles(int n, int granule, PackedMatrix* pm) {
    const int n16 = round_up(n, granule);
    pm->n = n;
    pm->n_tiles = ceil_div(n16, kMaxTileN);
    pm->tile_n = round_up(ceil_div(n16, pm->n_tiles), granule);
    pm->n_pad = pm->tile_n * pm->n_tiles;
}

// rows: function giving source row r (0..n-1) as (pointer to K0 floats, pointer to K1 floats) plus scale/bias
struct RowSource {
    const float* w0; const float* w1; double scale; double bias;
};

// `mode`: OperandMode of the stored weights (exact fp32, tf32-rounded fp32, fp16)
template <typename F>
int pack_matrix(Arena& arena, int n, int k0, int k1, int granule, int mode, bool with_bias, F row_of, PackedMatrix* pm) {
    const bool round = mode == OPERAND_TF32;
    const int chunk = mode == OPERAND_F16 ? kChunkKHalf : kChunkK;
    choose_tiles(n, granule, pm);
    pm->half = mode == OPERAND_F16 ? 1 : 0;
    pm->kseg[0] = k0; pm->kseg[1] = k1;
    pm->koff[0] = 0; pm->koff[1] = round_up(k0, chunk);
    pm->ld = round_up(k0, chunk) + (k1 > 0 ? round_up(k1, chunk) : 0);
    std::vector<float> hw((size_t)pm->n_pad * pm->ld, 0.0f), hb((size_t)pm->n_pad + 32, 0.0f);   // bias padded for vector loads
    for (int r = 0; r < n; ++r) {
        RowSource src = row_of(r);
        float* dst = &hw[(size_t)r * pm->ld];
        for (int k = 0; k < k0; ++k) dst[k] = (float)(src.scale * (double)src.w0[k]);
        for (int k = 0; k < k1; ++k) dst[pm->koff[1] + k] = (float)(src.scale * (double)src.w1[k]);
        if (round) for (int64_t k = 0; k < pm->ld; ++k) dst[k] = host_round_tf32(dst[k]);
        hb[r] = (float)src.bias;
    }
    if (mode == OPERAND_F16) {
        std::vector<__half> hh(hw.size());
        for (size_t i = 0; i < hw.size(); ++i) hh[i] = __float2half_rn(hw[i]);
        __half* d;
        EMPOSE_TRY(arena.upload(hh, &d));
        pm->w = reinterpret_cast<float*>(d);
    } else {
        EMPOSE_TRY(arena.upload(hw, &pm->w));
    }
    if (with_bias) EMPOSE_TRY(arena.upload(hb, &pm->bias));
    return EMPOSE_OK;
}

struct MlpPacked { std::vector<PackedMatrix> layers; };   // input, 2*blocks hidden, output

// nn.Linear (+ BatchNorm1d in eval mode folded in double) (+ PReLU slope) -> PackedMatrix
inline int pack_linear(Arena& arena, const TensorTable& tt, const std::string& lin, const std::string& bn,
                const std::string& prelu, int n_out, int n_in, int mode, PackedMatrix* pm) {
    const float *w, *b, *g = nullptr, *be = nullptr, *mu = nullptr, *var = nullptr;
    EMPOSE_TRY(tt.get_f32(lin + ".weight", {n_out, n_in}, &w));
    EMPOSE_TRY(tt.get_f32(lin + ".bias", {n_out}, &b));
    if (!bn.empty()) {
        EMPOSE_TRY(tt.get_f32(bn + ".weight", {n_out}, &g));
        EMPOSE_TRY(tt.get_f32(bn + ".bias", {n_out}, &be));
        EMPOSE_TRY(tt.get_f32(bn + ".running_mean", {n_out}, &mu));
        EMPOSE_TRY(tt.get_f32(bn + ".running_var", {n_out}, &var));
    }
    EMPOSE_TRY(pack_matrix(arena, n_out, n_in, 0, 16, mode, true, [&](int r) {
        RowSource s{w + (size_t)r * n_in, nullptr, 1.0, (double)b[r]};
        if (g) {   // y = gamma (Wx + b - mean) / sqrt(var + eps) + beta   (eps = 1e-5, torch default used by layers.py:26,57)
            const double sc = (double)g[r] / std::sqrt((double)var[r] + 1e-5);
            s.scale = sc;
            s.bias = ((double)b[r] - (double)mu[r]) * sc + (double)be[r];
        }
        return s;
    }, pm));
    if (!prelu.empty()) {
        const float* a;
        EMPOSE_TRY(tt.get_f32(prelu + ".weight", {1}, &a));
        pm->has_act = 1;
        pm->alpha = a[0];
    }
    return EMPOSE_OK;
}

// MLP of empose/nn/layers.py:46-77 with the reference's state-dict key layout
inline int pack_mlp(Arena& arena, const TensorTable& tt, const std::string& prefix, int n_in, int n_out, int hidden, int blocks,
             bool bn, int mode, MlpPacked* out) {
    out->layers.clear();
    out->layers.resize(2 + 2 * blocks);
    EMPOSE_TRY(pack_linear(arena, tt, prefix + ".input_to_hidden", bn ? prefix + ".batch_norm" : "", prefix + ".activation_fn",
                           hidden, n_in, mode, &out->layers[0]));
    const int stride = bn ? 4 : 3;
    for (int b = 0; b < blocks; ++b)
        for (int l = 0; l < 2; ++l) {
            const std::string base = prefix + ".hidden_layers." + std::to_string(b) + ".layers.";
            EMPOSE_TRY(pack_linear(arena, tt, base + std::to_string(l * stride), bn ? base + std::to_string(l * stride + 1) : "",
                                   base + std::to_string(l * stride + (bn ? 2 : 1)), hidden, hidden, mode,
                                   &out->layers[1 + 2 * b + l]));
        }
    EMPOSE_TRY(pack_linear(arena, tt, prefix + ".hidden_to_output", "", "", n_out, hidden, mode, &out->layers.back()));
    return EMPOSE_OK;
}

// ---- execution plan -------------------------------------------------------------------------------
struct JobRange { int begin = 0, count = 0, per_item = 1; };

struct MapKey {
    const void* ptr; int64_t stride; int k; int64_t rows; int box; int half;
    bool operator<(const MapKey& o) const {
        return std::tie(ptr, stride, k, rows, box, half) < std::tie(o.ptr, o.stride, o.k, o.rows, o.box, o.half);
    }
};

// one K segment of an A operand; strides / extents count elements (fp32, or fp16 when half = 1)
struct ASrc { const float* ptr = nullptr; int64_t stride = 0; int k = 0; int64_t rows = 0; int half = 0; };

// pointer to element `index` of an operand buffer that holds fp32 (half = 0) or fp16 (half = 1) elements
inline float* operand_at(float* base, size_t index, int half) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(base) + index * (half ? 2 : 4));
}
inline const float* operand_at(const float* base, size_t index, int half) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + index * (half ? 2 : 4));
}

struct JobBook {        // jobs + tensor maps of one plan
    bool use_tc = false;
    std::vector<GemmJob> jobs;
    std::vector<uint8_t> maps;     // kTensorMapBytes each
    std::map<MapKey, int> map_index;
    GemmJob* d_jobs = nullptr;
    void* d_maps = nullptr;

    int get_map(const float* ptr, int64_t stride, int k, int64_t rows, int box, int half, int* out) {
        *out = -1;
        if (!use_tc) return EMPOSE_OK;
        MapKey key{ptr, stride, k, rows, box, half};
        auto it = map_index.find(key);
        if (it != map_index.end()) { *out = it->second; return EMPOSE_OK; }
        const int idx = (int)(maps.size() / kTensorMapBytes);
        maps.resize(maps.size() + kTensorMapBytes);
        EMPOSE_TRY(tc_encode_map(&maps[(size_t)idx * kTensorMapBytes], ptr, stride, k, rows, box, half));
        map_index[key] = idx;
        *out = idx;
        return EMPOSE_OK;
    }

    // fp16 linear outputs the tcgen05 executor may write with TMA stores ([32 rows x 64 columns] boxes, gemm_tc.cu): needs
    // 16-byte aligned rows and column origin; rows >= m_rows are clipped by the map
    int attach_out_map(GemmJob& j) {
        j.out_map1 = 0;
        j.out2_map1 = 0;
        if (use_tc && j.epi == EPI_LINEAR && !j.out_half && !j.res && !j.has_act && !j.mask_rows && j.out) {
            // fp32 outputs of a plain contraction (the blend GEMMs): [32 rows x 32 columns] boxes; columns beyond the valid
            // ones are clipped by the map, so its extent is exactly what the job may write
            const int n_out = std::min(j.n_valid, j.split);
            if (!(reinterpret_cast<uintptr_t>(j.out) & 15) && !((j.out_stride * 4) & 15) && !(j.out_col0 & 3) && !(j.n_begin & 31) && n_out > 0) {
                int idx = -1;
                EMPOSE_TRY(get_map(j.out, j.out_stride, j.out_col0 + n_out, j.m_rows, 32, 0, &idx));
                j.out_map1 = idx + 1;
            }
            if (j.out2 && j.n_valid > j.split && !(reinterpret_cast<uintptr_t>(j.out2) & 15) && !((j.out2_stride * 4) & 15) && !(j.split & 31)) {
                int idx = -1;
                EMPOSE_TRY(get_map(j.out2, j.out2_stride, j.n_valid - j.split, j.m_rows, 32, 0, &idx));
                j.out2_map1 = idx + 1;
            }
            return EMPOSE_OK;
        }
        if (!use_tc || j.epi != EPI_LINEAR || !j.out_half || j.res || !j.out) return EMPOSE_OK;
        if ((reinterpret_cast<uintptr_t>(j.out) & 15) || ((j.out_stride * 2) & 15) || (j.out_col0 & 7) || (j.n_begin & 7)) return EMPOSE_OK;
        int idx = -1;
        EMPOSE_TRY(get_map(j.out, j.out_stride, (int)j.out_stride, j.m_rows, 32, 1, &idx));
        j.out_map1 = idx + 1;
        return EMPOSE_OK;
    }

    // fp16 LSTM jobs: a tensor map of the cell state [m_rows][hidden] fp32 (boxes of 32 rows x 32 units)
    int attach_c_map(GemmJob& j) {
        j.c_map1 = 0;
        if (!use_tc || j.epi != EPI_LSTM || !j.out_half || !j.c_state || j.gates_out || (j.hidden & 3)) return EMPOSE_OK;
        if (reinterpret_cast<uintptr_t>(j.c_state) & 15) return EMPOSE_OK;
        int idx = -1;
        EMPOSE_TRY(get_map(j.c_state, j.hidden, j.hidden, j.m_rows, 32, 0, &idx));
        j.c_map1 = idx + 1;
        // ... and of this step's hidden states [m_rows][hidden] fp16 (boxes of 32 rows x 32 units, 64-byte swizzle) in out_map1
        if (j.out && !(reinterpret_cast<uintptr_t>(j.out) & 15) && !((j.out_stride * 2) & 15) && !(j.hidden & 31)) {
            EMPOSE_TRY(get_map(j.out, j.out_stride, j.hidden, j.m_rows, 32, 2, &idx));
            j.out_map1 = idx + 1;
        }
        return EMPOSE_OK;
    }

    // appends one job per N tile of `W`; `proto` carries the epilogue fields (n_begin/n_count/maps are filled here)
    int add(const PackedMatrix& W, const ASrc& a0, const ASrc& a1, GemmJob proto, int m_rows, int dep, JobRange* range) {
        if (range->count == 0) range->begin = (int)jobs.size();
        if (range->begin + range->count != (int)jobs.size()) {      // a launch runs jobs [begin, begin + count)
            set_last_error("internal error: the jobs of a range must be added contiguously");
            return EMPOSE_E_ARG;
        }
        // `dep` counts jobs from the start of the range; ranges that use it run as ONE item per row tile (per_item == count),
        // so this is also the index inside the item.  The job it names must publish its stores to the TMA (gemm_tc.cu).
        if (dep >= 0) {
            int32_t& flag = jobs[(size_t)range->begin + dep].is_dep;
            const int kind = range->count - dep >= 2 ? 2 : 1;          // distance of the first job added here from the one it waits for
            flag = flag ? std::min(flag, kind) : kind;
        }
        for (int t = 0; t < W.n_tiles; ++t) {
            GemmJob j = proto;
            j.a_ptr[0] = a0.ptr; j.a_stride[0] = a0.stride; j.a_k[0] = a0.k;
            j.a_ptr[1] = a1.ptr; j.a_stride[1] = a1.stride; j.a_k[1] = a1.k;
            if (a0.half != W.half || (a1.k > 0 && a1.half != W.half)) {
                set_last_error("internal error: A and W operands of a job must have the same element type");
                return EMPOSE_E_ARG;
            }
            j.in_half = W.half;
            EMPOSE_TRY(get_map(a0.ptr, a0.stride, a0.k, a0.rows, kTileM, W.half, &j.a_map[0]));
            j.a_map[1] = -1;
            if (a1.k > 0) EMPOSE_TRY(get_map(a1.ptr, a1.stride, a1.k, a1.rows, kTileM, W.half, &j.a_map[1]));
            j.w_ptr = W.w; j.w_ld = W.ld; j.w_koff[0] = W.koff[0]; j.w_koff[1] = W.koff[1];
            EMPOSE_TRY(get_map(W.w, W.ld, (int)W.ld, W.n_pad, W.tile_n, W.half, &j.w_map));
            EMPOSE_TRY(get_map(W.w, W.ld, (int)W.ld, W.n_pad, W.tile_n / 2, W.half, &j.w_map2));
            j.n_begin = t * W.tile_n;
            j.n_count = W.tile_n;
            j.m_rows = m_rows;
            j.dep = dep;
            j.is_dep = 0;
            j.bias = W.bias;
            EMPOSE_TRY(attach_out_map(j));
            EMPOSE_TRY(attach_c_map(j));
            jobs.push_back(j);
            hot_dirty = true;
            ++range->count;
        }
        return EMPOSE_OK;
    }

    // the epilogue's one-load summary of a job (GemmJob::hot_*); is_dep is only final once every job has been added
    bool hot_dirty = false;
    void fill_hot() {
        for (GemmJob& j : jobs) {
            const bool plain = j.epi == EPI_LINEAR && j.out_half && !j.res && j.out_map1 > 0 &&
                               (j.out_scale == 0.0f || j.out_scale == 1.0f) && !j.mask_rows && !j.wait_ctr[0] && !j.wait_ctr[1] &&
                               !j.done_ctr && j.n_count == kMaxTileN && std::min(j.n_valid, j.split) - j.n_begin >= kMaxTileN;
            j.hot_path = plain ? (1 | (j.out_scratch ? 1 << 8 : 0) | ((j.is_dep & 3) << 9)) : 0;
            j.hot_out_col = j.out_col0 + j.n_begin;
            j.hot_out_map = j.out_map1 - 1;
            j.hot_alpha = j.has_act ? j.prelu_alpha : 1.0f;
        }
        hot_dirty = false;
    }

    int finalize(Arena& arena) {
        fill_hot();
        EMPOSE_TRY(arena.upload(jobs, &d_jobs));
        if (use_tc) {
            void* p;
            EMPOSE_TRY(arena.alloc(maps.size(), &p));
            EMPOSE_CUDA_TRY(cudaMemcpy(p, maps.data(), maps.size(), cudaMemcpyHostToDevice));
            d_maps = p;
        }
        return EMPOSE_OK;
    }
};

inline GemmJob linear_proto(const PackedMatrix& W, bool round, float* out, int64_t out_stride, int n_valid) {
    GemmJob j;
    memset(&j, 0, sizeof(j));
    j.epi = EPI_LINEAR;
    j.round_out = round ? 1 : 0;
    j.has_act = W.has_act;
    j.prelu_alpha = W.alpha;
    j.out_scale = W.out_scale;
    j.n_valid = n_valid;
    j.out = out;
    j.out_stride = out_stride;
    j.split = 1 << 30;
    j.frames_per_window = 1;
    return j;
}

// ---- the inference plan and the model context (model.cu) ------------------------------------------
struct Plan {
    int B = 0, F = 0, R = 0;
    uint64_t last_use = 0;      // plan-cache clock value of the call that used this plan last (LRU eviction)
    Arena arena;
    JobBook book;
    // workspace
    float *meas = nullptr, *xin = nullptr, *xiter = nullptr, *coef = nullptr;
    float *theta = nullptr, *beta = nullptr, *dtheta = nullptr, *dbeta = nullptr;
    float *pf = nullptr, *vpoff = nullptr, *dvp = nullptr, *dpf = nullptr, *gth_part = nullptr;
    float *jrest = nullptr, *dj = nullptr, *offsets = nullptr;     // rest joints / dE/dJ [R][kJrestLd]; packed offsets [B][12][12]
    float *joints = nullptr, *spos = nullptr, *sori = nullptr;
    int32_t* seq_len = nullptr;
    float* act[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [pose|shape][ping-pong]
    int64_t act_rows = 0;
    void* act_block = nullptr;     // the four activation buffers are one allocation (an L2 access-policy window can cover them)
    size_t act_block_bytes = 0;
    std::vector<float*> hseq, cstate, hinit;
    // staging for the host-buffer entry point
    float *in_pos = nullptr, *in_ori = nullptr, *in_masks = nullptr, *io_state = nullptr;
    float *in_off_r = nullptr, *in_off_t = nullptr;
    int32_t* in_len = nullptr;
    float *o_pose = nullptr, *o_shape = nullptr, *o_joints = nullptr;
    float* o_hist[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // the LSTM wavefront as ONE persistent launch (tc_launch_items): item table, per-(layer, step, row tile) completion
    // counters and the number of launches that have used them
    void* wave_items = nullptr;
    int wave_n_items = 0, wave_unit_rows = 0;
    uint32_t* wave_ctr = nullptr;
    uint32_t wave_epoch = 0;
    // job ranges
    std::vector<JobRange> lstm_diag;
    JobRange heads, init_chain, iter_chain, pb, pbt;
};


struct IefData {            // everything behind the opaque `empose_ief` handle
    empose_ief_config cfg;
    int num_sms = 148;
    int in_size = 0, iter_in = 0, n_pos = 0;
    int in_stride = 0, iter_stride = 0;   // row pitches of the network-input buffers (multiples of 4 floats for TMA)
    bool sensors_only = false; // made by empose_sensors_create: sub-model only, no learned layers
    bool round = true;         // a tensor-core mode (TF32 or FP16): pose-blend operands are tf32-rounded, tcgen05 executor
    int op_mode = OPERAND_TF32;   // storage of the MLP / LSTM / heads operands: OPERAND_F32, OPERAND_TF32 or OPERAND_F16
    int op_half = 0;           // op_mode == OPERAND_F16
    int pf_stride = kPoseFeatPad;   // elements per row of the feature buffer (2x when split hi|lo)
    bool blend_half = false;   // tensor-core modes: the blend GEMM and its transpose run on fp16 operands (feature rows
                               // split into fp16 hi | lo * 2^11, dE/dvp and dE/dJ stored as fp16 * kDvpScale); option blend_fp16 = 0
                               // keeps the tf32 form of round 1 (3xTF32 forward, single TF32 transposed)
    int dj_ld = kJrestLd;      // row pitch of the dE/dJ buffer in elements (kDjLdHalf when blend_half)
    Arena arena;
    SubModel sub;
    FanModel fan;              // fan tables of the sub-model (fan.ok = 0: the general kernel runs)
    ResidualSpec spec;
    int slot_of_sensor[kSensors];
    int static_tree = 0;       // sub.parents equals the standard SMPL body tree
    std::vector<PackedMatrix> lstm;
    PackedMatrix heads;
    MlpPacked pose_init, shape_init, pose_iter, shape_iter;
    PackedMatrix pb, pbt;
    std::map<std::tuple<int, int, int>, std::unique_ptr<Plan>> plans;     // (B, F, slot): slots > 0 serve the pipelined host path
    std::map<int, std::unique_ptr<Plan>> project_plans;
    // Plan cache policy: at most kMaxPlans / kMaxProjectPlans workspaces stay alive; the least recently used one is
    // evicted, one at a time, and never one that the current C-ABI call has already touched (its work may be in flight:
    // cudaFree would serialise the copy / compute overlap of the chunked host entry point).
    uint64_t plan_clock = 0;    // incremented per plan lookup
    uint64_t call_clock = 0;    // plan_clock at the start of the current C-ABI call
    int64_t last_launches = 0;
    // optional per-launch timing of the GEMM executor (empose_ief_set_profiling)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_main_events;    // same for the per-frame sub-model kernel
    size_t prof_main_used = 0;
    // copy streams and events of the pipelined host-buffer entry point (empose_ief_forward_host)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    std::vector<cudaEvent_t> pipe_events;
    cudaEvent_t slot_events[4][3] = {};      // streaming host entry point: per in-flight slot {uploaded, computed, downloaded}
    ~IefData() {
        for (auto& e : prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto& e : prof_main_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto& e : pipe_events) cudaEventDestroy(e);
        for (auto& se : slot_events) for (auto& e : se) if (e) cudaEventDestroy(e);
        if (copy_in) cudaStreamDestroy(copy_in);
        if (copy_out) cudaStreamDestroy(copy_out);
    }
};
}  // namespace empose

struct empose_ief : empose::IefData {};

namespace empose {
// uploads the "sub.*" arrays (SMPL sub-model, sensor topology) and packs the pose-blend operands into ctx->pb / ctx->pbt;
// needs ctx->round and ctx->arena (model.cu)
int upload_submodel(IefData* ctx, const TensorTable& tt);

// A operand of the blend GEMM: [pf_hi | pf_lo] then pf_hi again in TF32 mode, plain pf in FP32 mode
inline ASrc pose_blend_a0(const IefData* ctx, const float* pf, int rows) {
    return ASrc{pf, ctx->pf_stride, ctx->round ? 2 * kPoseFeatPad : kPoseFeatPad, rows, ctx->blend_half ? 1 : 0};
}
inline ASrc pose_blend_a1(const IefData* ctx, const float* pf, int rows) {
    return ctx->round ? ASrc{pf, ctx->pf_stride, kPoseFeatPad, rows, ctx->blend_half ? 1 : 0} : ASrc{};
}
// How the per-frame kernels store what feeds the blend GEMMs (UpdateParams::pf_mode, MainParams::round_out)
inline int blend_operand_mode(const IefData* ctx) { return !ctx->round ? OPERAND_F32 : ctx->blend_half ? OPERAND_F16 : OPERAND_TF32; }
// call once ctx->round is known (before upload_submodel)
inline void configure_blend(IefData* ctx) {
    ctx->blend_half = ctx->round && debug_options().blend_fp16 != 0;
    ctx->pf_stride = ctx->round ? 2 * kPoseFeatPad : kPoseFeatPad;
    ctx->dj_ld = ctx->blend_half ? kDjLdHalf : kJrestLd;
}

// The two contractions around the per-frame sub-model kernel, as jobs of `book`:
//   forward     [pf | beta] . [P ; S | 0 ; Jdirs] + [v_template | J0]  ->  vp [R][vp_dim], jrest [R][kJrestLd]
//   transposed  [dvp | dJ] . [P^T ; 0 | S^T ; Jdirs^T]                  ->  dpf [R][kPoseFeatPad] (dE/dpf | . | dE/dbeta)
// (reference: the shape blend, pose blend and joint regression of the BodyModel call at smpl.py:121 and their backward).
inline int add_blend_jobs(JobBook& book, const IefData* ctx, const float* pf, float* vp, float* jrest, int R, JobRange* fwd) {
    GemmJob proto = linear_proto(ctx->pb, false, vp, ctx->sub.vp_dim, ctx->sub.vp_dim + kPoseDim);
    proto.split = ctx->sub.vp_dim;
    proto.out2 = jrest;
    proto.out2_stride = kJrestLd;
    return book.add(ctx->pb, pose_blend_a0(ctx, pf, R), pose_blend_a1(ctx, pf, R), proto, R, -1, fwd);
}
inline int add_blend_transposed_jobs(JobBook& book, const IefData* ctx, const float* dvp, const float* dj, float* dpf, int R, JobRange* bwd) {
    GemmJob proto = linear_proto(ctx->pbt, false, dpf, kPoseFeatPad, kFeatK);
    const int hf = ctx->blend_half ? 1 : 0;
    return book.add(ctx->pbt, ASrc{dvp, ctx->sub.vp_dim, ctx->sub.vp_dim, R, hf}, ASrc{dj, ctx->dj_ld, kPoseDim, R, hf}, proto, R, -1, bwd);
}
}  // namespace empose
