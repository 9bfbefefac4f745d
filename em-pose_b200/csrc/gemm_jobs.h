// GEMM job descriptors and the epilogues shared by the two job executors
// (gemm_tc.cu: tcgen05 + TMA + TMEM, gemm_simt.cu: fp32 FFMA).
//
// A job is one output tile-column of one "layer":  D[M x n_count] = A[M x K] . W[n_begin.., K]^T
// followed by an epilogue that turns the fp32 accumulators into whatever the layer produces.
// A can be the concatenation of up to two K-segments taken from different buffers (the LSTM uses
// [h_{t-1} | x_t]).  Jobs are row-tile agnostic: executors apply a job to a 128-row tile at row m0.
//
// Reference semantics implemented by the epilogues:
//   EPI_LINEAR : nn.Linear (+ folded BatchNorm1d) + optional PReLU + optional LinearLayers skip
//                (empose/nn/layers.py:13-77); also the plain contraction of the pose blend.
//   EPI_LSTM   : one LSTM cell update with packed-sequence masking (empose/nn/layers.py:146-153).
#pragma once

#include <stdint.h>

#include "common.cuh"

namespace empose {

constexpr int kTileM = 128;       // rows per tile (UMMA M)
constexpr int kMaxTileN = 256;    // columns per job (UMMA N), multiple of 16
constexpr int kChunkK = 32;       // fp32 / tf32 elements per K chunk = one 128-byte swizzle row
constexpr int kChunkKHalf = 64;   // fp16 elements per K chunk (same 128 bytes)

enum EpilogueKind : int32_t { EPI_LINEAR = 0, EPI_LSTM = 1 };

struct GemmJob {
    // ---- what the tcgen05 executor's epilogue needs of the commonest job -- an fp16 linear layer on a full 256-column tile
    // ---- that leaves through TMA stores -- in ONE 16-byte load (JobBook::add fills it; 0 in `hot_path` = look at the fields) ----
    int32_t hot_path;        // bits 0-7: 1 = such a job; bit 8: out_scratch; bits 9-10: is_dep
    int32_t hot_out_col;     // out_col0 + n_begin
    int32_t hot_out_map;     // out_map1 - 1
    float hot_alpha;         // PReLU slope (1: no activation)
    // ---- A operand: up to two K segments ----
    const float* a_ptr[2];
    int64_t a_stride[2];     // floats between consecutive rows
    int32_t a_k[2];          // real K extent of the segment (0 = unused); padded to kChunkK in W
    int32_t a_map[2];        // tensor-map slot (tcgen05 executor)
    int32_t a_scratch[2];    // 1: the segment is CTA-local scratch: rows are indexed by 128 * blockIdx.x, not by the tile
    int32_t in_half;         // 1: A and W hold fp16 elements (kind::f16, 64-element K chunks); strides / extents count elements
    int32_t out_half;        // 1: `out` (and, for LSTM jobs, h_prev) hold fp16 elements
    int32_t out_scratch;     // 1: the output is CTA-local scratch (tcgen05 executor only)
    int32_t c_map1;          // 1 + tensor-map slot of `c_state` with a [32 rows x 32 fp32] box (tcgen05 executor, fp16 LSTM jobs: the cell-state
                             // block of an epilogue warp travels by TMA), 0: none
    int32_t out2_map1;       // the same for fp32 `out2` (boxes of [32 rows x 32 fp32]; fp32 plain contractions), 0: none
    int32_t out_map1;        // 1 + tensor-map slot of `out` with a [32 rows x 64 fp16] box (tcgen05 executor: the fp16 linear epilogue
                             // leaves through TMA stores), 0: none -- the epilogue stores from registers
    // ---- W operand: packed [n_total][w_ld], K-major, segment s starts at column w_koff[s] ----
    const float* w_ptr;
    int64_t w_ld;
    int32_t w_koff[2];
    int32_t w_map;           // tensor map with a box of n_count rows
    int32_t w_map2;          // tensor map with a box of n_count / 2 rows (2-CTA cluster variant: each CTA loads half)
    int32_t n_begin;         // first W row / output column of this job
    int32_t n_count;         // columns computed (multiple of 16, <= kMaxTileN)
    int32_t m_rows;          // valid rows overall; rows >= m_rows are never written
    int32_t dep;             // chain-local index of the job that must have finished before A may be loaded (-1: none)
    int32_t is_dep;          // != 0: a later job of the same item names this job as its `dep`: its epilogue must hand the
                             // stores of this thread (this job's and the earlier ones') over to the async proxy (TMA).
                             // 2: every such job comes at least two jobs later (the hand-over may be noticed lazily), 1: the next job may
    // ---- epilogue ----
    int32_t epi;
    int32_t round_out;       // 1: round outputs that feed later GEMMs to tf32 and use fast transcendentals
    const float* bias;       // indexed by global column (n_begin + c); may be null
    int32_t has_act;
    float prelu_alpha;
    float out_scale;         // accumulator scale applied before the bias (0 or 1: none); the fp16 blend GEMMs keep their
                             // operands in range with power-of-two factors and undo them here
    int32_t n_valid;         // columns >= n_valid (global index) are discarded
    float* out;              // columns [0, split)
    int64_t out_stride;
    int32_t out_col0;        // added to the global column index
    int32_t split;           // columns >= split go to out2 at (col - split)
    float* out2;
    int64_t out2_stride;
    const float* res;        // optional residual added after the activation (LinearLayers skip), same indexing as out
    int64_t res_stride;
    // row masking: rows whose frame index (row % frames_per_window) >= seq_len[row / frames_per_window]
    // contribute only the bias (the LSTM output of a padded frame is zero, layers.py:153)
    const int32_t* seq_len;
    int32_t frames_per_window;
    int32_t mask_rows;       // number of logical rows when rows are masked by sequence length (0: no masking)
    // ---- LSTM ----
    float* c_state;          // [M][H] cell state, updated in place
    const float* h_prev;     // carry source for frozen rows (row stride h_prev_stride)
    int64_t h_prev_stride;
    int32_t t;               // time step of this job
    int32_t hidden;          // H
    // training only (may be null): activated gates [row][4H] in the packed column order of this job (groups of 32 =
    // [i f g o] x 8 units) and the cell state after the step [row][H]
    float* gates_out;
    int64_t gates_stride;
    float* c_seq_out;
    int64_t c_seq_stride;
    // ---- cross-CTA ordering inside ONE launch (the persistent LSTM wavefront, tc_launch_items) ----
    // Counters are indexed by the 128-row tile; a job is complete for a tile when its epilogue has stored it.  A job may
    // start on a tile once the counters it names have reached wait_need * epoch (epoch = launches so far on this plan:
    // counters are never reset).  Null pointers: no ordering (the per-diagonal launches).
    const uint32_t* wait_ctr[2];
    uint32_t wait_need[2];
    uint32_t* done_ctr;
};

// Columns of an LSTM job are packed in groups of 32 = 8 hidden units x 4 gates:
// packed column n -> unit (n / 32) * 8 + n % 8, gate (n % 32) / 8  with gates ordered i, f, g, o.
__host__ __device__ inline int lstm_unit_of_packed(int n) { return (n / 32) * 8 + (n % 8); }
__host__ __device__ inline int lstm_gate_of_packed(int n) { return (n % 32) / 8; }

#if defined(__CUDACC__)
constexpr int kStageFloats = 32 * 32;              // one warp-private 4 KB staging tile: [32][32] fp32 with XOR-swizzled columns, or
                                                   // [32 rows][128 B] in the 128-byte TMA swizzle (1024-byte aligned in the tcgen05 executor)

// Epilogue of 32 consecutive accumulator columns [c0, c0+32) (c0 relative to the job, multiple of 32)
// for the 32 rows [row0, row0+32) owned by one warp: lane l holds row row0+l in `v` (the TMEM lane
// layout).  Results go through a warp-private shared-memory tile so that every global store (and the
// LSTM cell-state read-modify-write) is a fully coalesced row segment instead of 32 scattered words.
// Loads are batched into registers before any store so that pointer aliasing cannot serialise them.
// Bias arrays are padded by 32 floats, so the vector loads below never leave the allocation.
// Must be called by all 32 lanes of the warp.
// 16-byte accesses to the staging tiles in the shared state space proper (through a generic pointer they compile to
// LD.E / ST.E, whose latency sits on the long scoreboard)
__device__ __forceinline__ uint32_t smem_addr_of(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
    sts128(addr, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    const uint4 v = lds128(addr);
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}

// ---- fp16 linear jobs (the MLP chains) ----
// Everything the chunk loop needs of such a job, fetched ONCE per job into registers: the job record lives in memory the
// compiler must assume any store may alias, so fields read on demand are re-read after every store.
struct LinearHalfView {
    __half* out;             // first output element of the job's column 0 in row 0 (out_col0 and n_begin folded in)
    int64_t out_stride;
    const float* bias;       // bias of the job's column 0, or null
    float alpha;             // PReLU slope (1: no activation)
    int32_t m_rows;
    int32_t fast_cols;       // chunks with c0 + 32 <= fast_cols take this path (0: none)
    bool zero_row;           // this lane's row is a padded frame: it contributes only the bias (layers.py:153)
    int32_t out_map;         // tensor-map slot for TMA stores of [32 rows x 64 columns] blocks, -1: none
    int32_t out_col;         // column of the job's column 0 in that map
};
// `row0`: first OUTPUT row of the warp (a CTA-local scratch row for out_scratch jobs), `mask_row0`: its first LOGICAL row
__device__ __forceinline__ LinearHalfView linear_half_view(const GemmJob& j, int row0, int lane, int mask_row0) {
    LinearHalfView lv;
    const bool ok = j.epi == EPI_LINEAR && j.out_half && !j.res && (j.out_scale == 0.0f || j.out_scale == 1.0f);
    const int n_begin = j.n_begin;
    lv.fast_cols = ok ? min(j.n_valid, j.split) - n_begin : 0;
    lv.out = reinterpret_cast<__half*>(j.out) + j.out_col0 + n_begin;
    lv.out_stride = j.out_stride;
    lv.bias = j.bias ? j.bias + n_begin : nullptr;
    lv.alpha = j.has_act ? j.prelu_alpha : 1.0f;
    lv.m_rows = j.m_rows;
    lv.out_map = ok ? j.out_map1 - 1 : -1;
    lv.out_col = j.out_col0 + n_begin;
    const int row = mask_row0 + lane;
    lv.zero_row = false;
    if (ok && j.mask_rows && row < j.mask_rows) {
        const int fpw = j.frames_per_window;
        lv.zero_row = (row % fpw) >= j.seq_len[row / fpw];
    }
    return lv;
}
__device__ __forceinline__ void linear_half_load_bias(const LinearHalfView& lv, int c0, float (&bias)[32]) {
    if (lv.bias) {
        const float4* bp = reinterpret_cast<const float4*>(lv.bias + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 t = __ldg(bp + q);
            bias[4 * q] = t.x; bias[4 * q + 1] = t.y; bias[4 * q + 2] = t.z; bias[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bias[i] = 0.0f;
    }
}
// staged [32 rows][64 B] block (80-byte pitch) -> global memory, 8 rows per instruction
__device__ __forceinline__ void linear_half_flush(const LinearHalfView& lv, int row0, int lane, int c0, uint32_t tile, bool no_store) {
    __half* out_h = lv.out + c0;
    const int64_t out_stride = lv.out_stride;
    const int rows = no_store ? 0 : lv.m_rows - row0;
    uint4 val[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) val[it] = lds128(tile + (it * 8 + (lane >> 2)) * 80 + (lane & 3) * 16);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        if (r < rows) *reinterpret_cast<uint4*>(out_h + (int64_t)(row0 + r) * out_stride + (lane & 3) * 8) = val[it];
    }
    __syncwarp();
}
// A lane's 32 columns are 64 contiguous bytes of ITS row, so storing straight from registers makes every store
// instruction touch 32 different lines with 16 bytes each (measured: 40 % of a 512x512 layer).  Instead the warp stages
// its [32 rows][64 B] block in shared memory (80-byte pitch: conflict-free 16-byte accesses) and writes it out 8 rows
// per instruction: 4 lanes x 16 B = two full sectors per row.
// Step 1: bias, PReLU, fp16 -- 32 accumulator columns of this lane's row become 16 packed words
__device__ __forceinline__ void linear_half_pack(const LinearHalfView& lv, const float (&v)[32], const float (&bias)[32], uint32_t (&packed)[16]) {
    const float alpha = lv.alpha;
    const bool zero = lv.zero_row;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float y0 = (zero ? 0.0f : v[2 * i]) + bias[2 * i], y1 = (zero ? 0.0f : v[2 * i + 1]) + bias[2 * i + 1];
        y0 = y0 > 0.0f ? y0 : alpha * y0;
        y1 = y1 > 0.0f ? y1 : alpha * y1;
        const __half2 h = __floats2half2_rn(y0, y1);
        packed[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
}
// The same step for 64 columns with Blackwell's packed fp32 arithmetic (add / mul.f32x2 on register pairs) and PReLU as
// max(y, alpha y) for alpha <= 1, min(y, alpha y) otherwise (identical to y > 0 ? y : alpha y for every finite alpha, NaN and
// infinities propagate): 2.5 instructions per element instead of 4.5.
__device__ __forceinline__ void add_f32x2(float& x0, float& x1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b, d;\n\t"
        "mov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\t"
        "add.rn.f32x2 d, a, b;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}"
        : "+f"(x0), "+f"(x1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul_f32x2(float& t0, float& t1, float x0, float x1, float s) {
    asm("{\n\t.reg .b64 a, b, d;\n\t"
        "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\t"
        "mul.rn.f32x2 d, a, b;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}"
        : "=f"(t0), "=f"(t1) : "f"(x0), "f"(x1), "f"(s));
}
__device__ __forceinline__ void linear_half_pack64(const LinearHalfView& lv, float (&v)[64], const float (&bias)[64], uint32_t (&packed)[32]) {
    const float alpha = lv.alpha;
    if (__any_sync(0xffffffffu, lv.zero_row)) {        // padded frames (row-masked jobs only)
        if (lv.zero_row) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = 0.0f;
        }
    }
    if (alpha <= 1.0f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float t0, t1;
            add_f32x2(v[2 * i], v[2 * i + 1], bias[2 * i], bias[2 * i + 1]);
            mul_f32x2(t0, t1, v[2 * i], v[2 * i + 1], alpha);
            const __half2 h = __floats2half2_rn(fmaxf(v[2 * i], t0), fmaxf(v[2 * i + 1], t1));
            packed[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float t0, t1;
            add_f32x2(v[2 * i], v[2 * i + 1], bias[2 * i], bias[2 * i + 1]);
            mul_f32x2(t0, t1, v[2 * i], v[2 * i + 1], alpha);
            const __half2 h = __floats2half2_rn(fminf(v[2 * i], t0), fminf(v[2 * i + 1], t1));
            packed[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
    }
}
// The same with the bias read from a shared-memory copy eight values at a time (the tcgen05 executor keeps a second
// accumulator read in flight meanwhile: 64 live bias registers would spill).  No row masking (fast linear jobs have none).
__device__ __forceinline__ void linear_half_pack64_smem(float alpha, float (&v)[64], uint32_t bias_sa, int c0, uint32_t (&packed)[32]) {
    const bool use_max = alpha <= 1.0f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 b0 = lds128f(bias_sa + (uint32_t)(c0 + 8 * g) * 4u), b1 = lds128f(bias_sa + (uint32_t)(c0 + 8 * g + 4) * 4u);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float &x0 = v[8 * g + 2 * i], &x1 = v[8 * g + 2 * i + 1];
            float t0, t1;
            add_f32x2(x0, x1, bb[2 * i], bb[2 * i + 1]);
            mul_f32x2(t0, t1, x0, x1, alpha);
            const __half2 h = use_max ? __floats2half2_rn(fmaxf(x0, t0), fmaxf(x1, t1)) : __floats2half2_rn(fminf(x0, t0), fminf(x1, t1));
            packed[4 * g + i] = *reinterpret_cast<const uint32_t*>(&h);
        }
    }
}
// Step 2: through the staging tile to global memory
__device__ __forceinline__ void linear_half_store(const LinearHalfView& lv, int row0, int lane, int c0, const uint32_t (&packed)[16],
                                                  float* __restrict__ stage, bool no_store) {
    const uint32_t tile = smem_addr_of(stage);
#pragma unroll
    for (int q = 0; q < 4; ++q) sts128(tile + lane * 80 + q * 16, packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    __syncwarp();
    linear_half_flush(lv, row0, lane, c0, tile, no_store);
}
__device__ __forceinline__ void linear_half_chunk(const LinearHalfView& lv, int row0, int lane, int c0, const float (&v)[32],
                                                  const float (&bias)[32], float* __restrict__ stage, bool no_store) {
    uint32_t packed[16];
    linear_half_pack(lv, v, bias, packed);
    linear_half_store(lv, row0, lane, c0, packed, stage, no_store);
}

// Cell-state block of one chunk pair [c0, c0 + 64) of an fp16 LSTM job -> registers, in the lane mapping the tile fill of
// epilogue_chunk uses (8 rows x 64 B per instruction).  The tcgen05 executor issues this a pair ahead, so that the L2
// round trip overlaps the wait for the accumulator / the arithmetic of the previous pair.
__device__ __forceinline__ bool lstm_half_paired(const GemmJob& j) { return j.epi == EPI_LSTM && j.out_half && (j.n_count % 64) == 0; }
__device__ __forceinline__ void lstm_half_load_c(const GemmJob& j, int row0, int lane, int c0, float4 (&cpre)[4]) {
    const float* cs = j.c_state + lstm_unit_of_packed(j.n_begin + c0) + (lane & 3) * 4;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = row0 + it * 8 + (lane >> 2);
        cpre[it] = r < j.m_rows ? __ldcg(reinterpret_cast<const float4*>(cs + (int64_t)r * j.hidden)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// `bias_sa` != 0: shared-memory address of a copy of the job's bias (columns relative to the job; tcgen05 executor)
__device__ __forceinline__ void epilogue_chunk(const GemmJob& j, int row0, int lane, int c0, float (&v)[32],
                                               float* __restrict__ stage, bool no_store = false, const float4* cpre = nullptr,
                                               uint32_t bias_sa = 0, int mask_row0 = -1) {
    const int n0 = j.n_begin + c0;                 // global column of v[0]
    const int row = row0 + lane;
    const int mrow = (mask_row0 >= 0 ? mask_row0 : row0) + lane;      // logical row (differs from `row` for CTA-local scratch outputs)
    float bias[32];
    if (bias_sa) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 t = lds128f(bias_sa + (uint32_t)(c0 + 4 * q) * 4u);
            bias[4 * q] = t.x; bias[4 * q + 1] = t.y; bias[4 * q + 2] = t.z; bias[4 * q + 3] = t.w;
        }
    } else if (j.bias) {
        const float4* bp = reinterpret_cast<const float4*>(j.bias + n0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 t = __ldg(bp + q);
            bias[4 * q] = t.x; bias[4 * q + 1] = t.y; bias[4 * q + 2] = t.z; bias[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bias[i] = 0.0f;
    }
    if (j.epi == EPI_LINEAR && j.out_half && !j.res && n0 + 32 <= j.n_valid && n0 + 32 <= j.split &&
        (j.out_scale == 0.0f || j.out_scale == 1.0f)) {
        // fp16 activations (the tcgen05 executor builds the view once per job and calls linear_half_chunk directly)
        const LinearHalfView lv = linear_half_view(j, row0, lane, mask_row0 >= 0 ? mask_row0 : row0);
        linear_half_chunk(lv, row0, lane, c0, v, bias, stage, no_store);
    } else if (j.epi == EPI_LSTM && j.out_half && (j.n_count % 64) == 0) {
        // (job fields into registers first: after any store the compiler has to assume the job record changed)
        const int m_rows = j.m_rows, hidden = j.hidden, t_step = j.t;
        float* const c_state = j.c_state;
        float* const out = j.out;
        const float* const h_prev = j.h_prev;
        const int64_t out_stride = j.out_stride, h_prev_stride = j.h_prev_stride;
        const int32_t* const seq_len = j.seq_len;
        // fp16 hidden state, two chunks (16 units) at a time: the cell state of the pair is 64 B per row, the hidden state
        // 32 B.  They are moved between global and shared memory with coalesced 16-byte accesses (8 / 16 rows per
        // instruction, whole sectors) and every lane works on its own row in shared memory in between.
        const int pos = (c0 >> 5) & 1;                          // first / second chunk of the pair
        const int unit0 = lstm_unit_of_packed(n0) - pos * 8;    // first unit of the pair
        const uint32_t ctile = smem_addr_of(stage);             // [32][80 B]: 16 fp32 cell states per row
        const uint32_t htile = ctile + 32 * 80;                 // [32][48 B]: 16 fp16 hidden states per row
        const bool live = row < m_rows && t_step < seq_len[row];
        if (pos == 0) {
            const bool any_frozen = __any_sync(0xffffffffu, row < m_rows && !live);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = it * 8 + (lane >> 2), seg = lane & 3;
                if (cpre)
                    sts128f(ctile + r * 80 + seg * 16, cpre[it]);
                else if (row0 + r < m_rows)
                    sts128f(ctile + r * 80 + seg * 16,
                            __ldcg(reinterpret_cast<const float4*>(c_state + (int64_t)(row0 + r) * hidden + unit0 + seg * 4)));
            }
            if (any_frozen) {
                const __half* hprev = reinterpret_cast<const __half*>(h_prev) + unit0;
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int r = it * 16 + (lane >> 1), seg = lane & 1;
                    if (row0 + r < m_rows) {
                        const uint4 hv = __ldcg(reinterpret_cast<const uint4*>(hprev + (int64_t)(row0 + r) * h_prev_stride + seg * 8));
                        sts128(htile + r * 48 + seg * 16, hv.x, hv.y, hv.z, hv.w);
                    }
                }
            }
            __syncwarp();
        }
        if (live) {
            const uint32_t cp = ctile + lane * 80 + pos * 32;
            const float4 c0v = lds128f(cp), c1v = lds128f(cp + 16);
            const float c_old[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
            float c_new[8];
            uint32_t hp[4];
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                float h2[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float gi = v[k + q] + bias[k + q], gf = v[8 + k + q] + bias[8 + k + q];
                    const float gg = v[16 + k + q] + bias[16 + k + q], go = v[24 + k + q] + bias[24 + k + q];
                    c_new[k + q] = sigmoid_f(gf) * c_old[k + q] + sigmoid_f(gi) * tanh_f(gg);
                    h2[q] = sigmoid_f(go) * tanh_f(c_new[k + q]);
                }
                const __half2 h = __floats2half2_rn(h2[0], h2[1]);
                hp[k / 2] = *reinterpret_cast<const uint32_t*>(&h);
            }
            sts128f(cp, make_float4(c_new[0], c_new[1], c_new[2], c_new[3]));
            sts128f(cp + 16, make_float4(c_new[4], c_new[5], c_new[6], c_new[7]));
            sts128(htile + lane * 48 + pos * 16, hp[0], hp[1], hp[2], hp[3]);
        }
        if (pos == 1) {       // padded rows carry c (unchanged in the tile) and h (h_prev, loaded above)
            __syncwarp();
            float4 cv[4];
            uint4 hv[2];
#pragma unroll
            for (int it = 0; it < 4; ++it) cv[it] = lds128f(ctile + (it * 8 + (lane >> 2)) * 80 + (lane & 3) * 16);
#pragma unroll
            for (int it = 0; it < 2; ++it) hv[it] = lds128(htile + (it * 16 + (lane >> 1)) * 48 + (lane & 1) * 16);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = it * 8 + (lane >> 2), seg = lane & 3;
                if (row0 + r < m_rows)
                    *reinterpret_cast<float4*>(c_state + (int64_t)(row0 + r) * hidden + unit0 + seg * 4) = cv[it];
            }
            __half* hout = reinterpret_cast<__half*>(out) + unit0;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int r = it * 16 + (lane >> 1), seg = lane & 1;
                if (row0 + r < m_rows)
                    *reinterpret_cast<uint4*>(hout + (int64_t)(row0 + r) * out_stride + seg * 8) = hv[it];
            }
            __syncwarp();
        }
    } else if (j.epi == EPI_LSTM && j.out_half) {
        // fp16 hidden state, narrow layers (4H not a multiple of 64): a lane owns 8 consecutive units of its row
        const int unit0 = lstm_unit_of_packed(n0);
        if (row < j.m_rows) {
            __half* hout = reinterpret_cast<__half*>(j.out) + (int64_t)row * j.out_stride + unit0;
            if (j.t < j.seq_len[row]) {
                float4* cp = reinterpret_cast<float4*>(j.c_state + (int64_t)row * j.hidden + unit0);
                const float4 c0v = __ldcg(cp), c1v = __ldcg(cp + 1);
                const float c_old[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
                float c_new[8];
                uint32_t hp[4];
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    float h2[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float gi = v[k + q] + bias[k + q], gf = v[8 + k + q] + bias[8 + k + q];
                        const float gg = v[16 + k + q] + bias[16 + k + q], go = v[24 + k + q] + bias[24 + k + q];
                        c_new[k + q] = sigmoid_f(gf) * c_old[k + q] + sigmoid_f(gi) * tanh_f(gg);
                        h2[q] = sigmoid_f(go) * tanh_f(c_new[k + q]);
                    }
                    const __half2 h = __floats2half2_rn(h2[0], h2[1]);
                    hp[k / 2] = *reinterpret_cast<const uint32_t*>(&h);
                }
                cp[0] = make_float4(c_new[0], c_new[1], c_new[2], c_new[3]);
                cp[1] = make_float4(c_new[4], c_new[5], c_new[6], c_new[7]);
                *reinterpret_cast<uint4*>(hout) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            } else {      // padded step: the state is carried (packed-sequence semantics); c stays where it is
                const __half* hprev = reinterpret_cast<const __half*>(j.h_prev) + (int64_t)row * j.h_prev_stride + unit0;
                *reinterpret_cast<uint4*>(hout) = __ldcg(reinterpret_cast<const uint4*>(hprev));
            }
        }
    } else if (j.epi == EPI_LINEAR && !j.out_half && !j.res && !j.has_act && !j.mask_rows && n0 + 32 <= j.n_valid &&
               (n0 >= j.split || n0 + 32 <= j.split) &&
               ((n0 >= j.split ? ((n0 - j.split) | (int)j.out2_stride) : ((j.out_col0 + n0) | (int)j.out_stride)) & 3) == 0) {
        // fp32 outputs of a plain contraction (the blend GEMMs), a whole chunk on one side of `split`, 16-byte aligned rows:
        // two passes of 16 columns through the staging tile ([32 rows][64 B], 80-byte pitch: conflict-free 128-bit
        // accesses), written out 8 rows x 64 B per instruction -- 24 memory instructions per chunk instead of 96.
        const float scale = j.out_scale != 0.0f ? j.out_scale : 1.0f;
        const int round_out = j.round_out;
        float* dst;
        int64_t stride;
        if (n0 >= j.split) { dst = j.out2 + (n0 - j.split); stride = j.out2_stride; }
        else { dst = j.out + j.out_col0 + n0; stride = j.out_stride; }
        const int rows = no_store ? 0 : j.m_rows - row0;
        const uint32_t tile = smem_addr_of(stage);
#pragma unroll
        for (int hcol = 0; hcol < 32; hcol += 16) {
            float y[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                y[i] = fmaf(v[hcol + i], scale, bias[hcol + i]);
                if (round_out) y[i] = round_tf32(y[i]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) sts128f(tile + lane * 80 + q * 16, make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]));
            __syncwarp();
            float4 val[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) val[it] = lds128f(tile + (it * 8 + (lane >> 2)) * 80 + (lane & 3) * 16);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = it * 8 + (lane >> 2);
                if (r < rows) *reinterpret_cast<float4*>(dst + (int64_t)(row0 + r) * stride + hcol + (lane & 3) * 4) = val[it];
            }
            __syncwarp();
        }
    } else if (j.epi == EPI_LINEAR) {
        if (j.mask_rows && mrow < j.mask_rows && (mrow % j.frames_per_window) >= j.seq_len[mrow / j.frames_per_window]) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.0f;
        }
        const float alpha = j.has_act ? j.prelu_alpha : 1.0f;
        const float scale = j.out_scale != 0.0f ? j.out_scale : 1.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float y = fmaf(v[i], scale, bias[i]);
            y = y > 0.0f ? y : alpha * y;
            v[i] = y;
        }
        if (j.res && row < j.m_rows) {
            const float* rp = j.res + (int64_t)row * j.res_stride + j.out_col0 + n0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < j.n_valid) v[i] += rp[i];
        }
        if (j.round_out) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = round_tf32(v[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[lane * 32 + (i ^ lane)] = v[i];      // column i of row `lane` lives at i ^ lane: conflict-free both ways
        __syncwarp();
        const int n = n0 + lane;
        const bool col_ok = (c0 + lane < j.n_count) && (n < j.n_valid);
        float* dst;
        int stride;
        if (n < j.split) { dst = j.out + j.out_col0 + n; stride = (int)j.out_stride; }
        else { dst = j.out2 + (n - j.split); stride = (int)j.out2_stride; }
        const int rows = col_ok ? min(32, j.m_rows - row0) : 0;
        if (j.out_half) {                              // fp16 output on an edge chunk (rare): element-wise stores
            __half* hd = reinterpret_cast<__half*>(j.out) + j.out_col0 + n + (int64_t)row0 * stride;
            for (int r = 0; r < rows; ++r) hd[(int64_t)r * stride] = __float2half_rn(stage[r * 32 + (lane ^ r)]);
        } else {
            dst += (int64_t)row0 * stride;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {          // batches of 8: loads first, then stores, few live registers
                float o[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) o[r] = stage[(r0 + r) * 32 + (lane ^ (r0 + r))];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r0 + r < rows) dst[(r0 + r) * stride] = o[r];
            }
        }
        __syncwarp();
    } else {  // EPI_LSTM: 8 hidden units x 4 gates per chunk
        const int unit0 = lstm_unit_of_packed(n0);
        float* sc = stage;                 // [32][9] cell state
        float* sh = stage + 32 * 9;        // [32][9] carried / new hidden state
        const int sub = lane >> 3, u = lane & 7;
        const int rows = j.m_rows - row0;
        float* cg = j.c_state + (int64_t)row0 * j.hidden + unit0 + u;
        const float* hg = j.h_prev + (int64_t)row0 * j.h_prev_stride + unit0 + u;
        const bool live = row < j.m_rows && j.t < j.seq_len[row];
        // the carried hidden state is only needed for rows whose sequence has ended (packed-sequence semantics)
        const bool any_frozen = __any_sync(0xffffffffu, row < j.m_rows && !live);
        float cl[8], hl[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {      // coalesced: 4 rows x 8 units per instruction
            const int r = q * 4 + sub;
            cl[q] = r < rows ? __ldcg(cg + (int64_t)r * j.hidden) : 0.0f;
        }
        if (any_frozen) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = q * 4 + sub;
                hl[q] = r < rows ? __ldcg(hg + (int64_t)r * j.h_prev_stride) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) sh[(q * 4 + sub) * 9 + u] = hl[q];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) sc[(q * 4 + sub) * 9 + u] = cl[q];
        __syncwarp();
        if (live) {
            // new cell / hidden state go straight back into the staging tiles; the activated gates replace the
            // pre-activations in v[] (no second register array: this path must not cost the other paths registers)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float gi = v[k] + bias[k];
                const float gf = v[8 + k] + bias[8 + k];
                const float gg = v[16 + k] + bias[16 + k];
                const float go = v[24 + k] + bias[24 + k];
                const float c_old = sc[lane * 9 + k];
                float c_new, h_new, si, sf, tg, so;
                if (j.round_out) {
                    si = sigmoid_f(gi); sf = sigmoid_f(gf); tg = tanh_f(gg); so = sigmoid_f(go);
                    c_new = sf * c_old + si * tg;
                    h_new = round_tf32(so * tanh_f(c_new));
                } else {
                    si = 1.0f / (1.0f + expf(-gi)); sf = 1.0f / (1.0f + expf(-gf)); tg = tanhf(gg); so = 1.0f / (1.0f + expf(-go));
                    c_new = sf * c_old + si * tg;
                    h_new = so * tanhf(c_new);
                }
                v[k] = si; v[8 + k] = sf; v[16 + k] = tg; v[24 + k] = so;
                sc[lane * 9 + k] = c_new;
                sh[lane * 9 + k] = h_new;
            }
            if (j.gates_out) {
                // training: keep the activated gates for the backward pass, in the accumulator's own column order
                // (32 contiguous floats per row and chunk: eight 16-byte stores instead of 32 scattered words)
                float4* gp = reinterpret_cast<float4*>(j.gates_out + (int64_t)row * j.gates_stride + n0);
#pragma unroll
                for (int q = 0; q < 8; ++q) gp[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            cl[q] = sc[(q * 4 + sub) * 9 + u];
            hl[q] = sh[(q * 4 + sub) * 9 + u];
        }
        float* og = j.out + (int64_t)row0 * j.out_stride + unit0 + u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = q * 4 + sub;
            if (r < rows) {
                cg[(int64_t)r * j.hidden] = cl[q];
                og[(int64_t)r * j.out_stride] = hl[q];
                if (j.c_seq_out) j.c_seq_out[(int64_t)(row0 + r) * j.c_seq_stride + unit0 + u] = cl[q];
            }
        }
        __syncwarp();
    }
}
#endif

}  // namespace empose
