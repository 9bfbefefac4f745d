// tcgen05 / TMA / TMEM job executor (sm_100a).
//
// One persistent, warp-specialised kernel executes lists of GemmJob (gemm_jobs.h) on 128-row tiles:
//
//   warp 0   TMA producer   cp.async.bulk.tensor.2d (128B swizzle) of A [128 x 128 B] and W [n x 128 B] K chunks (32 fp32 /
//                           tf32 or 64 fp16 elements) into a shared-memory ring (3 slots of 48 KB, or 5 of 32 KB in CTA-pair
//                           mode), completion on mbarriers
//   warp 1   MMA issuer     tcgen05.mma kind::f16 / kind::tf32 (M=128 per CTA, N<=256, 32 bytes of K per instruction),
//                           accumulating in TMEM; tcgen05.commit releases ring slots and publishes accumulators.  In the
//                           default mode pairs of CTAs form ONE cta_group::2 MMA (M=256), issued by CTA 0 of the pair.
//                           Both roles run their loops on the whole warp with uniform values; an elect.sync lane issues.
//   warp 2   TMEM allocator 512 columns = two 128x256 fp32 accumulators (double buffered across jobs)
//   warp 3   job stager     copies the record and the bias of the NEXT job into one of two shared-memory slots (mbarriers
//                           job_full / job_empty), so the epilogue never reads job fields or biases from global memory
//   warps 4-11 epilogue     two warps per TMEM lane quadrant, each taking 128 of the 256 columns, every warp on its own: it
//                           waits for the job slot and the accumulator, reads 64 columns per tcgen05.ld (two reads software
//                           pipelined), and its output leaves through TMA stores from two private 4 KB staging tiles (fp16
//                           and fp32 linear jobs; LSTM jobs: the cell-state block comes and goes by TMA, the hidden state
//                           goes by TMA); it hands the accumulator back itself (tmem_empty counts the warps) and publishes
//                           what later jobs wait for.  EMPOSE_TC_TRACE writes a per-job timeline of the three roles.
//
// A work item is (row tile, group of consecutive jobs).  Jobs of one item run back to back in the
// same CTA; a job may depend on an earlier job of its item (an MLP layer reading the previous layer's
// activations): the producer then waits for that job's epilogue, whose global stores are ordered at CTA
// scope and handed to the TMA (async proxy) with fence.proxy.async before it is counted as done.  Independent jobs in
// between (the pose and shape MLPs are interleaved layer by layer) keep the tensor pipe busy meanwhile.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../include/empose_b200.h"
#include "gemm_jobs.h"
#include "gemm_tc.h"

namespace empose {

namespace {

#ifndef EMPOSE_EPI_WARPS
#define EMPOSE_EPI_WARPS 8
#endif
constexpr int kEpiWarps = EMPOSE_EPI_WARPS;  // 8 or 12: kEpiParts warps per TMEM lane quadrant share the columns of an accumulator.
                                             // Measured (profiles/r02/README.md): twelve warps with 32-column chunks (56 / 152
                                             // registers, a 5-slot ring) run the MLP chain no faster than eight warps with
                                             // 64-column chunks (810 vs 806 us) and the LSTM wavefront slower (1.33 vs 1.19 ms): the
                                             // chain is not bound by epilogue parallelism.  Eight is the default.
constexpr int kEpiParts = kEpiWarps / 4;
static_assert(kEpiWarps == 8 || kEpiWarps == 12, "two or three epilogue warps per TMEM lane quadrant");
constexpr int kDefaultClusterMode = 2;      // EMPOSE_TC_CLUSTER when the variable is not set: CTA pairs (cta_group::2)
constexpr int kABytes = kTileM * kChunkK * 4;        // 16 KB
constexpr int kWBytes = kMaxTileN * kChunkK * 4;     // 32 KB
constexpr int kStageBytes = kABytes + kWBytes;
// operand ring: 192 KB with 8 epilogue warps, 160 KB with 12 (their staging tiles need the difference)
constexpr int kStages = 3;                           // ring slots of A [128 x 128 B] + W [256 x 128 B] (48 KB)
constexpr int kStagesPair = 5;                       // CTA-pair mode: a slot holds half of W (32 KB), so the same memory is a deeper ring
                                                     // (6 -> 4 slots cost the MLP chain nothing and a lone GEMM 4 %: the sixth slot's 32 KB
                                                     //  are better spent on a second staging tile per epilogue warp)
constexpr int kOperandBytes = kStagesPair * (kABytes + kWBytes / 2);
static_assert(kStages * kStageBytes <= kOperandBytes, "both ring layouts share the operand region");
constexpr int kTmemCols = 512;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kEpiThreads = kEpiWarps * 32;

struct __align__(8) Control {
    uint64_t full[kStagesPair];
    uint64_t empty[kStagesPair];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t job_full[2];             // the stager warp has put a job record and its bias into slot b
    uint64_t job_empty[2];            // every epilogue warp is done with slot b
    uint32_t tmem_base;
    uint64_t c_bar[kEpiWarps];        // per epilogue warp: its cell-state block has arrived in its staging tile (LSTM jobs)
    volatile uint32_t done_seq[kEpiWarps];   // per epilogue warp: 1 + sequence number of the last job whose stores it has handed to the TMA
    uint32_t job_done_cnt[4];         // per job (sequence number & 3): epilogue warps that have published their part (cross-CTA jobs)
    int32_t job_unit[2];              // item-table launches: the row-tile unit of the job in slot b (written by the stager warp)
    volatile uint32_t dep_seq;        // 1 + sequence number of the last job whose cross-CTA counters the producer warp has seen satisfied
};

constexpr int kEpiTiles = 2;                                     // staging tiles per epilogue warp: TMA stores alternate between them
constexpr int kEpiStageBytes = kEpiWarps * kEpiTiles * kStageFloats * 4;     // 4 KB tiles, 1024-byte aligned (TMA swizzle)
static_assert(kOperandBytes % 1024 == 0 && (kStageFloats * 4) % 1024 == 0, "staging tiles must keep the 1024-byte alignment of the swizzle pattern");
constexpr int kBiasBytes = 2 * kMaxTileN * 4;                    // the bias of the current and the next job
constexpr int kJobSlotBytes = ((int)sizeof(GemmJob) + 15) / 16 * 16;
constexpr int kControlBytes = 288;
constexpr int kJobWords = (int)(sizeof(GemmJob) / 4);
static_assert(sizeof(GemmJob) % 8 == 0 && kJobWords <= 3 * 32, "the stager warp copies a job record with three words per lane");
static_assert(sizeof(Control) <= kControlBytes, "Control grew");
// dynamic shared memory, in this order; the base is 1024-byte aligned (checked at kernel start)
constexpr int kOffStage = kOperandBytes, kOffBias = kOffStage + kEpiStageBytes, kOffJobs = kOffBias + kBiasBytes,
              kOffControl = kOffJobs + 2 * kJobSlotBytes;
constexpr int kSmemBytes = kOffControl + kControlBytes;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// What the single-thread roles need of a job.  They fetch the fields of the NEXT job while working on the current one:
// every epilogue ends in a device-scope fence, which invalidates L1, so a field read on demand is an L2 round trip on
// the critical path (ncu, v10: 38 % of all stall samples were waits for such loads).
struct ProducerView {
    int32_t dep, n_count, n_begin, in_half, a_k[2], a_map[2], a_scratch[2], w_map, w_map2, w_koff[2];
    const uint32_t* wait_ctr[2];
    uint32_t wait_need[2];
};
struct IssuerView {
    int32_t in_half, n_count, a_k[2];
};
__device__ __forceinline__ ProducerView producer_view(const GemmJob& j) {
    ProducerView v;
    v.dep = j.dep; v.n_count = j.n_count; v.n_begin = j.n_begin; v.in_half = j.in_half;
    v.a_k[0] = j.a_k[0]; v.a_k[1] = j.a_k[1]; v.a_map[0] = j.a_map[0]; v.a_map[1] = j.a_map[1];
    v.a_scratch[0] = j.a_scratch[0]; v.a_scratch[1] = j.a_scratch[1];
    v.w_map = j.w_map; v.w_map2 = j.w_map2; v.w_koff[0] = j.w_koff[0]; v.w_koff[1] = j.w_koff[1];
    v.wait_ctr[0] = j.wait_ctr[0]; v.wait_ctr[1] = j.wait_ctr[1]; v.wait_need[0] = j.wait_need[0]; v.wait_need[1] = j.wait_need[1];
    return v;
}
__device__ __forceinline__ IssuerView issuer_view(const GemmJob& j) {
    IssuerView v;
    v.in_half = j.in_half; v.n_count = j.n_count; v.a_k[0] = j.a_k[0]; v.a_k[1] = j.a_k[1];
    return v;
}

// ---- timeline trace (EMPOSE_TC_TRACE=<file>): clock64 stamps of the three roles per job, first CTAs / jobs of a launch ----
constexpr int kTraceCtas = 4, kTraceJobs = 96, kTraceStamps = 4, kTraceRoles = 4;       // role 3: the epilogue warp's job start in detail
__device__ unsigned long long* g_trace_buf = nullptr;       // [cta][role 0 producer, 1 issuer, 2 epilogue warp 4][job][stamp]
// `t` = g_trace_buf read ONCE per role (a stamp must not wait for a load: after a fence that is an L2 round trip)
__device__ __forceinline__ void trace_stamp(unsigned long long* t, int role, uint32_t seq, int k) {
    if (t && seq < kTraceJobs) t[((blockIdx.x * kTraceRoles + role) * kTraceJobs + seq) * kTraceStamps + k] = clock64();
}

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int crd0, int crd1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(crd0), "r"(crd1)
        : "memory");
}

// multicast variants: the tile / the arrival lands at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int crd0, int crd1,
                                                      uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(crd0), "r"(crd1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int crd0, int crd1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(crd0), "r"(crd1), "l"(policy)
        : "memory");
}
// CTA-pair (cta_group::2) variants.  The tile lands in the executing CTA's shared memory, the bytes are counted on a
// barrier that may live in the peer CTA (`bar_cluster_addr` is a shared::cluster address, see mapa_rank0).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int crd0, int crd1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(crd0), "r"(crd1)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// shared::cluster address of the same shared-memory location in CTA 0 of the cluster
__device__ __forceinline__ uint32_t mapa_rank0(const void* p) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
    return r;
}
// Arrival on a barrier that may live in the peer CTA.  Default semantics (release at CTA scope) on purpose: with
// .release.cluster ptxas emits MEMBAR.ALL.GPU + ERRBAR in front of the arrive, which stalled the arriving thread -- and with it
// the whole epilogue, which meets at a barrier every job -- for ~3500-5400 cycles per job (timeline trace, profiles/r02): the
// thread had to wait for its outstanding global / TMA stores.  What the arrival orders here are tensor-memory reads, which
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync have already completed.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// spin with a watchdog: a protocol bug must end in a trap (an error the host sees), never in a hung GPU
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}

// ---- cross-CTA ordering inside one launch (GemmJob::wait_ctr / done_ctr) ----
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin (with back-off and a watchdog that traps) until a completion counter has reached `need`
__device__ __forceinline__ void wait_counter(const uint32_t* p, uint32_t need) {
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_gpu(p) - need) < 0) {
        __nanosleep(40);
        if (++spins > (1u << 25)) __trap();
    }
}

// two counters at once: both acquire loads are in flight together (each is an L2 round trip); null pointers are skipped
__device__ __forceinline__ void wait_counters2(const uint32_t* p0, uint32_t need0, const uint32_t* p1, uint32_t need1, int tile) {
    uint32_t spins = 0;
    for (;;) {
        const uint32_t v0 = p0 ? ld_acquire_gpu(p0 + tile) : need0;
        const uint32_t v1 = p1 ? ld_acquire_gpu(p1 + tile) : need1;
        if ((int32_t)(v0 - need0) >= 0 && (int32_t)(v1 - need1) >= 0) return;
        __nanosleep(40);
        if (++spins > (1u << 25)) __trap();
    }
}

// One lane of a fully active, converged warp.  The TMA and MMA roles run their loops on the WHOLE warp with uniform
// values and let the elected lane issue: inside an `if (lane == 0)` region the compiler treats every operand as divergent
// and wraps each UTMALDG / UTCHMMA in ELECT + R2UR shuffles -- ~200 instructions per K chunk for four MMAs, which made the
// issuing thread, not the tensor pipe, the mainloop's bound (ncu: pipe 58 % active with nothing else running).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- TMA stores (epilogue): a staged [32 rows x 128 B] block in the 128-byte swizzle -> global memory, rows beyond the tensor clipped ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int crd0, int crd1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(crd0), "r"(crd1)
                 : "memory");
}
// L2 eviction priorities: CTA-local scratch activations (written by a layer's epilogue, read back by the next layer's loads,
// overwritten by the next tile) should stay in L2 -- evict_last -- instead of being written back to HBM under the pressure of
// the streaming operands; they are 75 MB per launch of the 126 MB.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t smem_src, int crd0, int crd1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(crd0), "r"(crd1), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }     // sources may be overwritten
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }     // ... all but the latest store's
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }           // the writes are complete
// orders this thread's async-proxy accesses to GLOBAL memory behind what it has acquired (the recurrent state another CTA wrote
// through its TMA); bit 262144 of the debug mask brings the all-spaces fence back for comparison
__device__ __forceinline__ void fence_async_global(int debug_mode) {
    if (debug_mode & 262144) asm volatile("fence.proxy.async;" ::: "memory");
    else asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// CTA-pair MMA: D is [256 x N]: rows 0-127 accumulate in the TMEM of CTA 0, rows 128-255 in CTA 1's; each CTA's shared
// memory holds ITS 128 rows of A and ITS N/2 rows of W at the offsets the descriptors name.  Issued by CTA 0 only.
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor for a K-major tile whose rows are 128 bytes (32 fp32) apart, stored
// with the 128-byte swizzle TMA produces: 8-row groups are 1024 bytes apart (SBO), LBO is unused.
// Field layout: cute::UMMA::SmemDescriptor (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1
// [46,48), layout_type [61,64) with SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10) or f16 (0, 0), both
// K-major, N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc(int n, int half, int m = kTileM) {
    const uint32_t fmt = half ? 0u : ((2u << 7) | (2u << 10));
    return (1u << 4) | fmt | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_load_32cols(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 64 accumulator columns in ONE instruction: half as many waits per column as two x32 reads
__device__ __forceinline__ void tmem_load_64cols(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_load_32cols_nowait(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr));
}
// the same read without the wait: the bias is fetched from shared memory while the accumulator is on its way
__device__ __forceinline__ void tmem_load_64cols_nowait(uint32_t taddr, float (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31]), "=f"(v[32]),
          "=f"(v[33]), "=f"(v[34]), "=f"(v[35]), "=f"(v[36]), "=f"(v[37]), "=f"(v[38]), "=f"(v[39]), "=f"(v[40]),
          "=f"(v[41]), "=f"(v[42]), "=f"(v[43]), "=f"(v[44]), "=f"(v[45]), "=f"(v[46]), "=f"(v[47]), "=f"(v[48]),
          "=f"(v[49]), "=f"(v[50]), "=f"(v[51]), "=f"(v[52]), "=f"(v[53]), "=f"(v[54]), "=f"(v[55]), "=f"(v[56]),
          "=f"(v[57]), "=f"(v[58]), "=f"(v[59]), "=f"(v[60]), "=f"(v[61]), "=f"(v[62]), "=f"(v[63])
        : "r"(taddr));
}
// (the values of an unwaited read must not be touched before this; "memory" keeps the compiler from moving their uses up)
__device__ __forceinline__ void tmem_load_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// N bias values of the job's columns [c0, c0 + N) from the stager warp's shared-memory copy (the same address in every lane: broadcast)
template <int N>
__device__ __forceinline__ void bias_from_smem(uint32_t bias_sa, int c0, float (&b)[N]) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
        const float4 t = lds128f(bias_sa + (uint32_t)(c0 + 4 * q) * 4u);
        b[4 * q] = t.x; b[4 * q + 1] = t.y; b[4 * q + 2] = t.z; b[4 * q + 3] = t.w;
    }
}

// ---- fp16 LSTM jobs, the fast epilogue: one warp, 32 rows x 128 gate columns = 32 hidden units per row ----
// NaN-propagating minimum (a clamp must not swallow the non-finite values the precision guard of the host class looks for)
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// One LSTM cell (layers.py:146-153 -> torch.nn.LSTM) from the four gate pre-activations and the old cell state with SEVEN
// special-function operations instead of ten: with e_i = exp(-x_i), e_f = exp(-x_f), e_g = exp(2 x_g), e_o = exp(-x_o)
//   c' = sigmoid(x_f) c + sigmoid(x_i) tanh(x_g) = [c (1 + e_i)(1 + e_g) + (e_g - 1)(1 + e_f)] / [(1 + e_i)(1 + e_g)(1 + e_f)]
//   h  = sigmoid(x_o) tanh(c')                   = (e_c - 1) / [(1 + e_c)(1 + e_o)],   e_c = exp(2 c')
// The exponents are clamped where the functions have long reached their limits in fp32 (sigmoid(-20) = 2e-9, tanh(15) = 1 - 2e-13),
// which keeps every product below 3e30.  The special-function pipe (4 lanes per cycle and scheduler) was the floor of the old
// epilogue: 10 operations x 8192 cells per tile = 5120 cycles of the ~5900 the MMAs of a tile take.
__device__ __forceinline__ void lstm_cell_fast(float xi, float xf, float xg, float xo, float c_old, float& c_new, float& h) {
    constexpr float kL2e = 1.4426950408889634f;
    const float ei = ex2_approx(min_nan(xi * -kL2e, 20.0f * kL2e));
    const float ef = ex2_approx(min_nan(xf * -kL2e, 20.0f * kL2e));
    const float eg = ex2_approx(min_nan(xg * (2.0f * kL2e), 30.0f * kL2e));
    const float eo = ex2_approx(min_nan(xo * -kL2e, 20.0f * kL2e));
    const float f1 = 1.0f + ef;
    const float p = (1.0f + ei) * (1.0f + eg);
    const float num = fmaf(c_old, p, (eg - 1.0f) * f1);
    c_new = num * rcp_approx(p * f1);
    const float ec = ex2_approx(min_nan(c_new * (2.0f * kL2e), 30.0f * kL2e));
    h = (ec - 1.0f) * rcp_approx((1.0f + ec) * (1.0f + eo));
}

// Register budgets of the warp roles (setmaxnreg): the four single-thread / idle warps hand registers to the eight epilogue
// warps, whose 64-column chunks keep the accumulator read, the bias and two packed halves live.  168 * 384 = 72 * 128 + 216 * 256; 128 * 512 = 56 * 128 + 152 * 384.
constexpr int kRegsControl = kEpiWarps == 8 ? 72 : 56;
constexpr int kRegsEpilogue = kEpiWarps == 8 ? 216 : 152;
constexpr bool kWideChunks = kEpiWarps == 8;       // the 64-column fp16 linear path needs ~190 registers

template <class View>
__device__ __forceinline__ int job_chunk_k(const View& j) { return j.in_half ? kChunkKHalf : kChunkK; }
template <class View>
__device__ __forceinline__ int job_k_chunks(const View& j) {
    const int ck = job_chunk_k(j);
    return (j.a_k[0] + ck - 1) / ck + (j.a_k[1] + ck - 1) / ck;
}

// ---- the kernel -----------------------------------------------------------------------------------
// kMode = 1: every CTA loads its own A and W tiles.
// kMode = 2: launched as clusters of two CTAs that walk the SAME job sequence on two adjacent row tiles.  Each CTA
//   loads its own A tile and ONE HALF of every W tile, multicast into both CTAs' shared memory, which halves the W
//   traffic out of L2.  A ring slot may only be refilled when BOTH CTAs have consumed it, so every MMA commit arrives
//   on the `empty` barrier of both CTAs.  (Measured: no faster -- L2 is not the bound, shared memory is.)
// kMode = 3: the same pairs of row tiles as ONE tcgen05.mma.cta_group::2 (M = 256): each CTA keeps its A tile and HALF
//   of the W tile in its own shared memory (32 KB per ring slot instead of 48) and the tensor cores of the two SMs
//   share the W halves.  CTA 0 issues the MMAs for both; its `full` barriers count the TMA bytes of both CTAs; MMA
//   commits arrive on both CTAs' `empty` / `tmem_full` barriers; both epilogues arrive on CTA 0's `tmem_empty`.
template <int kMode>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const GemmJob* __restrict__ jobs,
                                                              const CUtensorMap* __restrict__ maps, int job_begin,
                                                              int job_count, int jobs_per_item, int m_tiles,
                                                              int debug_mode, const int2* __restrict__ items, int n_items_table,
                                                              uint32_t epoch) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* epi_stage = reinterpret_cast<float*>(smem + kOffStage);
    float* bias_s = reinterpret_cast<float*>(smem + kOffBias);
    Control* ctl = reinterpret_cast<Control*>(smem + kOffControl);
    auto job_slot = [&](uint32_t b) -> GemmJob& { return *reinterpret_cast<GemmJob*>(smem + kOffJobs + b * kJobSlotBytes); };
    if (smem_u32(smem) & 1023u) __trap();      // the TMA swizzle patterns assume it

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int groups = job_count / jobs_per_item;
    // the job sequence of this CTA: (item, jj) -> index into `jobs`; the successor of the last job of an item is the first of the next
    // With an item table (the persistent LSTM wavefront) item i is the single job items[i].x on row-tile unit items[i].y, and
    // jobs order themselves across CTAs through completion counters (GemmJob::wait_ctr / done_ctr).
    auto job_index = [&](int item, int jj) { return items ? items[item].x : job_begin + (item % groups) * jobs_per_item + jj; };
    auto item_unit = [&](int item) { return items ? items[item].y : item / groups; };
    // work items: (row tile, job group); in cluster mode an item is a PAIR of row tiles, one per CTA of the cluster
    constexpr int kCluster = kMode >= 2 ? 2 : 1;
    constexpr bool kPair = kMode == 3;
    constexpr int kRing = kPair ? kStagesPair : kStages;                                    // slots in the operand ring ...
    constexpr int kSlotBytes = kPair ? kABytes + kWBytes / 2 : kStageBytes;                 // ... of this many bytes
    static_assert(kRing * kSlotBytes <= kOperandBytes, "the ring must fit the operand region");
    const uint32_t ring = (kPair && (debug_mode & 256)) ? (uint32_t)kStages : (uint32_t)kRing;     // bit 256: experiment, shallow ring
    const int n_items = items ? n_items_table : (kCluster == 2 ? (m_tiles + 1) / 2 : m_tiles) * groups;
    const uint32_t crank = kCluster == 2 ? cluster_ctarank() : 0u;
    const int item0 = kCluster == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = kCluster == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr uint16_t kMask = kCluster == 2 ? 3 : 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRing; ++s) {
            mbar_init(&ctl->full[s], 1);
            mbar_init(&ctl->empty[s], kMode == 2 ? 2 : 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&ctl->tmem_full[b], 1);
            mbar_init(&ctl->tmem_empty[b], (kPair ? 2 : 1) * kEpiWarps);      // every epilogue warp hands its lanes of the accumulator back
            mbar_init(&ctl->job_full[b], 1);
            mbar_init(&ctl->job_empty[b], kEpiWarps);
        }
        for (int w = 0; w < kEpiWarps; ++w) mbar_init(&ctl->c_bar[w], 1);
        for (int w = 0; w < kEpiWarps; ++w) ctl->done_seq[w] = 0;
        for (int q = 0; q < 4; ++q) ctl->job_done_cnt[q] = 0;
        ctl->dep_seq = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (kPair) {      // the same warp of both CTAs allocates the same columns in both tensor memories
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                         "n"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                         "n"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (kCluster == 2) cluster_sync_all();        // the peer's barriers must be initialised before anything is multicast to it
    tcgen05_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    unsigned long long* const trace = blockIdx.x < kTraceCtas ? g_trace_buf : nullptr;

    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));
        // ================= TMA producer (whole warp, one elected lane issues) =================
        {
            uint32_t stage = 0, phase = 0, items_done = 0, pseq = 0;
            const uint64_t keep_policy = l2_policy_evict_last();
            ProducerView nxt;
            if (item0 < n_items) nxt = producer_view(jobs[job_index(item0, 0)]);
            for (int item = item0; item < n_items; item += item_step, ++items_done) {
                const int m0 = (kCluster == 2 ? 2 * item_unit(item) + (int)crank : item_unit(item)) * kTileM;
                for (int jj = 0; jj < jobs_per_item; ++jj) {
                    const ProducerView job = nxt;
                    bool has_nxt = false;
                    {
                        const int nj = jj + 1 < jobs_per_item ? jj + 1 : 0;
                        const int ni = nj ? item : item + item_step;
                        has_nxt = ni < n_items;
                        if (has_nxt) nxt = producer_view(jobs[job_index(ni, nj)]);
                    }
                    if (lane == 0) trace_stamp(trace, 0, pseq, 0);
                    if (job.dep >= 0) {
                        // every epilogue warp has handed its part of that job's output to the async proxy
                        const uint32_t need = items_done * (uint32_t)jobs_per_item + (uint32_t)job.dep + 1u;
                        uint32_t spins = 0;
                        while (!__all_sync(0xffffffffu, lane >= kEpiWarps || ctl->done_seq[lane] >= need)) {
                            if (++spins > (1u << 28)) __trap();
                        }
                        __threadfence_block();
                        __syncwarp();
                    }
                    if (job.wait_ctr[0] || job.wait_ctr[1]) {
                        // this tile's A rows are produced by other CTAs of THIS launch: wait for their epilogues, then
                        // order the TMA reads (async proxy) after what the acquire made visible
                        if (m0 < m_tiles * kTileM)
                            wait_counters2(job.wait_ctr[0], job.wait_need[0] * epoch, job.wait_ctr[1], job.wait_need[1] * epoch, m0 / kTileM);
                        // the epilogue warps need the same counters (they read the recurrent state): they wait for this word in
                        // shared memory instead of polling the counters themselves -- an L2 round trip per job and warp
                        __threadfence_block();
                        if (lane == 0) ctl->dep_seq = pseq + 1u;
                        fence_async_global(debug_mode);
                        __syncwarp();
                    }
                    if (lane == 0) trace_stamp(trace, 0, pseq, 1);
                    const uint32_t w_bytes = (uint32_t)job.n_count * kChunkK * 4u;        // 128 bytes per row in either type
                    const int ck = job_chunk_k(job);
                    const int half_rows = job.n_count >> 1;
                    for (int seg = 0; seg < 2; ++seg) {
                        const int chunks = (job.a_k[seg] + ck - 1) / ck;
                        const CUtensorMap* a_map = &maps[job.a_map[seg]];
                        const int a_row = job.a_scratch[seg] ? (int)blockIdx.x * kTileM : m0;
                        const int w_k0 = job.w_koff[seg];
                        for (int kc = 0; kc < chunks; ++kc) {
                            // The NEXT job's tensor maps into the TMA unit's descriptor cache while this job's chunks stream: a
                            // job's first load otherwise starts with a descriptor fetch from L2 (every job of a chain names
                            // other maps; the timeline showed ~1200 cycles until a job's first operands arrived).
                            if (seg == 0 && kc == 1 && has_nxt && !(debug_mode & 131072)) {
                                if (elect_one()) {
                                    prefetch_tensormap(&maps[nxt.a_map[0]]);
                                    if (nxt.a_k[1] > 0) prefetch_tensormap(&maps[nxt.a_map[1]]);
                                    prefetch_tensormap(&maps[kPair || kCluster == 2 ? nxt.w_map2 : nxt.w_map]);
                                }
                                __syncwarp();
                            }
                            if (kCluster == 2) mbar_wait_guarded(&ctl->empty[stage], phase ^ 1u);
                            else mbar_wait(&ctl->empty[stage], phase ^ 1u);
                            uint8_t* a_dst = smem + stage * kSlotBytes;
                            uint8_t* w_dst = a_dst + kABytes;
                            if (elect_one()) {
                                if (kPair) {
                                    // both CTAs fill their own slot; the bytes of both are counted on CTA 0's barrier
                                    const uint32_t full0 = mapa_rank0(&ctl->full[stage]);
                                    if (crank == 0) mbar_arrive_expect_tx(&ctl->full[stage], 2u * (uint32_t)kABytes + w_bytes);
                                    if (job.a_scratch[seg] && !(debug_mode & 8192)) tma_load_2d_pair_hint(a_dst, a_map, full0, kc * ck, a_row, keep_policy);
                                    else tma_load_2d_pair(a_dst, a_map, full0, kc * ck, a_row);
                                    tma_load_2d_pair(w_dst, &maps[job.w_map2], full0, w_k0 + kc * ck, job.n_begin + (int)crank * half_rows);
                                } else if (debug_mode & 2) {            // measurement only: no loads, MMAs run on stale data
                                    mbar_arrive(&ctl->full[stage]);
                                } else {
                                    mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)kABytes + w_bytes);
                                    tma_load_2d(a_dst, a_map, &ctl->full[stage], kc * ck, a_row);
                                    if (kCluster == 2) {       // my half of the W rows, delivered to both CTAs
                                        tma_load_2d_multicast(w_dst + crank * (uint32_t)half_rows * 128u, &maps[job.w_map2], &ctl->full[stage],
                                                              w_k0 + kc * ck, job.n_begin + (int)crank * half_rows, kMask);
                                    } else {
                                        tma_load_2d(w_dst, &maps[job.w_map], &ctl->full[stage], w_k0 + kc * ck, job.n_begin);
                                    }
                                }
                            }
                            __syncwarp();
                            if (++stage == ring) { stage = 0; phase ^= 1u; }
                        }
                    }
                    if (lane == 0) trace_stamp(trace, 0, pseq, 2);
                    ++pseq;
                }
            }
        }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));
        // ================= MMA issuer (whole warp, one elected lane issues) =================
        if (!(kPair && crank != 0)) {
            uint32_t stage = 0, phase = 0, seq = 0;
            IssuerView nxt;
            if (item0 < n_items) nxt = issuer_view(jobs[job_index(item0, 0)]);
            const uint32_t ring_base = smem_u32(smem);
            for (int item = item0; item < n_items; item += item_step) {
                for (int jj = 0; jj < jobs_per_item; ++jj, ++seq) {
                    const IssuerView job = nxt;
                    {
                        const int nj = jj + 1 < jobs_per_item ? jj + 1 : 0;
                        const int ni = nj ? item : item + item_step;
                        if (ni < n_items) nxt = issuer_view(jobs[job_index(ni, nj)]);
                    }
                    const uint32_t buf = seq & 1u;
                    if (lane == 0) trace_stamp(trace, 1, seq, 0);
                    if (kCluster == 2) mbar_wait_guarded(&ctl->tmem_empty[buf], ((seq >> 1) & 1u) ^ 1u);
                    else mbar_wait(&ctl->tmem_empty[buf], ((seq >> 1) & 1u) ^ 1u);
                    tcgen05_fence_after();
                    if (lane == 0) trace_stamp(trace, 1, seq, 1);
                    const uint32_t d_tmem = tmem_base + buf * kMaxTileN;
                    const bool half = job.in_half != 0;
                    const uint32_t idesc = make_idesc(job.n_count, half, kPair ? 2 * kTileM : kTileM);
                    const int chunks = job_k_chunks(job);
                    for (int kc = 0; kc < chunks; ++kc) {
                        if (kCluster == 2) mbar_wait_guarded(&ctl->full[stage], phase);
                        else mbar_wait(&ctl->full[stage], phase);
                        tcgen05_fence_after();
                        if (kc == 0 && lane == 0) trace_stamp(trace, 1, seq, 2);
                        const uint64_t a_desc = make_smem_desc(ring_base + stage * (uint32_t)kSlotBytes);
                        const uint64_t b_desc = a_desc + (uint64_t)(kABytes >> 4);
                        if (elect_one()) {
                            if (debug_mode & 1) {                // measurement only: loads without MMAs (not in cluster mode)
                                mbar_arrive(&ctl->empty[stage]);
                            } else {
                                // 8 tf32 / 16 f16 = 32 bytes along K inside the 128-byte swizzle row per MMA: +2 in the >>4 address field
                                if (half) {
#pragma unroll
                                    for (int k = 0; k < kChunkK / 8; ++k) {
                                        if (kPair) umma_f16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
                                        else umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < kChunkK / 8; ++k) {
                                        if (kPair) umma_tf32_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
                                        else umma_tf32(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
                                    }
                                }
                                if (kPair) umma_commit_pair(&ctl->empty[stage], kMask);                     // frees the slot in both CTAs
                                else if (kCluster == 2) umma_commit_multicast(&ctl->empty[stage], kMask);   // ditto
                                else umma_commit(&ctl->empty[stage]);
                            }
                        }
                        __syncwarp();
                        if (++stage == ring) { stage = 0; phase ^= 1u; }
                    }
                    if (elect_one()) {
                        if (kPair) umma_commit_pair(&ctl->tmem_full[buf], kMask);     // both CTAs' epilogues read their half of D
                        else umma_commit(&ctl->tmem_full[buf]);
                    }
                    if (lane == 0) trace_stamp(trace, 1, seq, 3);
                    __syncwarp();
                }
            }
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));       // the TMEM allocator has nothing else to do
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));
        // ================= job stager =================
        // Keeps the epilogue warps supplied with the record and the bias of the NEXT job in shared memory (two slots), so that
        // their chunk loops never wait for global memory and need no CTA-wide barrier between jobs: every epilogue warp runs
        // on its own, paced only by the accumulator barriers.
        uint32_t seq = 0;
        for (int item = item0; item < n_items; item += item_step) {
            for (int jj = 0; jj < jobs_per_item; ++jj, ++seq) {
                const uint32_t b = seq & 1u;
                const GemmJob* src = &jobs[job_index(item, jj)];
                uint32_t w[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) w[q] = lane + 32 * q < kJobWords ? reinterpret_cast<const uint32_t*>(src)[lane + 32 * q] : 0u;
                const float* bias = src->bias;
                const int n_begin = src->n_begin, n4 = src->n_count >> 2;
                if (lane == 0 && !(debug_mode & 131072)) {      // the tensor maps the epilogue of that job will store (and load) through
                    const int m0i = src->out_map1, m1i = src->out2_map1, m2i = src->c_map1;
                    if (m0i > 0) prefetch_tensormap(&maps[m0i - 1]);
                    if (m1i > 0) prefetch_tensormap(&maps[m1i - 1]);
                    if (m2i > 0) prefetch_tensormap(&maps[m2i - 1]);
                }
                float4 bv[2];
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    bv[q] = (bias && lane + 32 * q < n4) ? __ldg(reinterpret_cast<const float4*>(bias + n_begin) + lane + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                mbar_wait(&ctl->job_empty[b], ((seq >> 1) & 1u) ^ 1u);
                const int item_unit_next = items ? items[item].y : 0;       // item-table launches: the row-tile unit goes along
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (lane + 32 * q < kJobWords) reinterpret_cast<uint32_t*>(&job_slot(b))[lane + 32 * q] = w[q];
#pragma unroll
                for (int q = 0; q < 2; ++q) reinterpret_cast<float4*>(bias_s + b * kMaxTileN)[lane + 32 * q] = bv[q];
                if (lane == 0) ctl->job_unit[b] = item_unit_next;
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->job_full[b]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
        // ================= epilogue =================
        const int ew = warp - 4;
        const int quad = ew & 3;                  // TMEM lanes 32*quad .. 32*quad+31 (a warp may only touch its own quadrant)
        const int part = ew >> 2;                 // which share of the accumulator's columns (kEpiParts warps per quadrant)
        uint32_t seq = 0;
        bool generic_stores = false;              // this thread has stored from registers since its last proxy fence
        bool tma_pending = false;                 // (warp-uniform) a TMA store of this warp may still be reading its staging tile
        const uint64_t keep_policy = l2_policy_evict_last();
        uint32_t c_phase = 0;                     // (warp-uniform) parity of the next completion of this warp's cell-state barrier
        uint32_t pub_pending = 0;                 // (warp-uniform) 1 + sequence number of a job whose TMA stores are not yet known to be complete
        // Wait until the staging tile may be rewritten.  A deferred publication rides on the same wait: by now the stores of the
        // job it belongs to are a chunk's worth of work old, so waiting for their completion instead of just their reads
        // costs next to nothing, whereas at the end of the job it was ~1000 cycles on every epilogue warp's critical path.
        auto release_tile = [&]() {
            if (tma_pending || pub_pending) {
                if (lane == 0) {
                    if (pub_pending) { bulk_wait_all0(); ctl->done_seq[ew] = pub_pending; }
                    else bulk_wait_read0();
                }
                __syncwarp();
                tma_pending = false;
                pub_pending = 0;
            }
        };
        // TMA stores of the linear paths alternate between the warp's two staging tiles: a block only has to wait for the store
        // before the previous one (in practice never), not for the one just issued.  Returns the tile's shared-memory address.
        uint32_t tile_sel = 0;
        auto next_store_tile = [&](float* stage0) -> uint32_t {
            if (lane == 0) {
                if (pub_pending) { bulk_wait_all0(); ctl->done_seq[ew] = pub_pending; }
                else if (tma_pending) bulk_wait_read1();
            }
            __syncwarp();
            pub_pending = 0;
            const uint32_t t = smem_addr_of(stage0) + tile_sel * (uint32_t)(kStageFloats * 4);
            tile_sel ^= 1u;
            return t;
        };
        const int et = (int)threadIdx.x - 4 * 32;
        for (int item = item0; item < n_items; item += item_step) {
            const int unit_direct = items ? 0 : item / groups;
            for (int jj = 0; jj < jobs_per_item; ++jj, ++seq) {
                const uint32_t buf = seq & 1u;
                if (et == 0) trace_stamp(trace, 3, seq, 0);
                mbar_wait(&ctl->job_full[buf], (seq >> 1) & 1u);
                if (et == 0) trace_stamp(trace, 3, seq, 1);
                const GemmJob& job = job_slot(buf);
                // item-table launches: the stager warp left the item's row-tile unit in the slot (a lookup in the table here was
                // an L2 round trip per job on the epilogue's serial path)
                const int unit = items ? ctl->job_unit[buf] : unit_direct;
                const int m0 = (kCluster == 2 ? 2 * unit + (int)crank : unit) * kTileM;
                const uint32_t bias_sa = smem_u32(bias_s + buf * kMaxTileN);
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * kMaxTileN;
                // The commonest job -- an fp16 linear layer on a full tile, out through TMA stores -- is summarised in the first
                // 16 bytes of its record: one load instead of ~20 dependent ones and their branches before the accumulator wait
                // (the timeline showed ~1000 cycles between two jobs of an epilogue warp, which is what bounds the MLP chain).
                const uint4 hot = lds128(smem_u32(&job));
                const bool fast_linear = (hot.x & 0xffu) == 1u && kWideChunks && !(debug_mode & (1024 | 2048 | 16384));
                // activations chained inside this CTA live in CTA-local scratch rows: they are re-read from L2 by the next
                // layer and overwritten by the next tile before they would be written back to HBM
                const bool out_scratch = fast_linear ? ((hot.x >> 8) & 1u) != 0u : job.out_scratch != 0;
                const int row0 = (out_scratch ? (int)blockIdx.x * kTileM : m0) + quad * 32;
                // column shares in units of 32-column chunks -- of chunk PAIRS for fp16 LSTM jobs, whose cell-state blocks span 64
                int c_begin, c_end;
                bool lstm_pre = false;
                if (fast_linear) {
                    c_begin = part * (kMaxTileN / kEpiParts);
                    c_end = c_begin + kMaxTileN / kEpiParts;
                } else {
                    const int unit = lstm_half_paired(job) ? 64 : 32;
                    const int units = (job.n_count + unit - 1) / unit;
                    c_begin = (part * units / kEpiParts) * unit;
                    c_end = min(job.n_count, ((part + 1) * units / kEpiParts) * unit);
                    // fp16 LSTM jobs: the cell state of the first chunk pair is fetched while the MMAs are still running
                    lstm_pre = lstm_half_paired(job) && !(debug_mode & 128);
                    // Wavefront jobs: the recurrent state this epilogue reads (cell state, carried hidden state) is written by
                    // other CTAs of this launch -- wait for them here too (the MMAs cannot start earlier either).  Both
                    // counters are polled at once (an acquire load is an L2 round trip).
                    if (job.wait_ctr[0] || job.wait_ctr[1]) {
                        uint32_t spins = 0;
                        while ((int32_t)(ctl->dep_seq - (seq + 1u)) < 0) {
                            if (++spins > (1u << 28)) __trap();
                        }
                        __threadfence_block();
                    }
                }
                // The fast LSTM epilogue (a full 256-column tile: this warp owns 128 gate columns = 32 hidden units of its 32 rows):
                // the warp's [32 rows x 32 units] fp32 cell-state block comes by TMA straight into its staging tile (128-byte
                // swizzle, conflict-free per-row reads) while the MMAs are still running, is updated in place and leaves by TMA.
                if (et == 0) trace_stamp(trace, 3, seq, 2);      // job fields read, cross-CTA dependencies met
                const bool lstm_fast = lstm_pre && job.c_map1 > 0 && c_end - c_begin == 128 && !job.gates_out && !(debug_mode & 4096);      // (false for fast_linear)
                float* my_stage = epi_stage + ew * (kEpiTiles * kStageFloats);
                float4 cpre[4];
                int lf_seq_len = 0;
                if (lstm_fast) {
                    release_tile();
                    if (lane == 0) {
                        // (the block was written by another CTA of this launch through ITS TMA; the counter wait above was the
                        //  acquire, this orders the async-proxy read behind it)
                        fence_async_global(debug_mode);
                        mbar_arrive_expect_tx(&ctl->c_bar[ew], 32u * 128u);
                        tma_load_2d(my_stage, &maps[job.c_map1 - 1], &ctl->c_bar[ew], lstm_unit_of_packed(job.n_begin + c_begin), row0);
                    }
                    if (row0 + lane < job.m_rows) lf_seq_len = __ldcg(job.seq_len + row0 + lane);
                } else if (lstm_pre && c_begin < c_end) {
                    lstm_half_load_c(job, row0, lane, c_begin, cpre);
                }
                if (et == 0) trace_stamp(trace, 3, seq, 3);      // recurrent state requested (LSTM jobs)
                // fp16 linear jobs: all fields the chunk loop needs, once per job
                LinearHalfView lv;
                bool f32_tma = false;
                float f32_scale = 1.0f;
                if (fast_linear) {
                    lv.out = nullptr; lv.out_stride = 0; lv.bias = nullptr; lv.m_rows = 0;        // (TMA stores clip; the bias is in shared memory)
                    lv.alpha = __uint_as_float(hot.w); lv.fast_cols = kMaxTileN; lv.zero_row = false;
                    lv.out_map = (int32_t)hot.z; lv.out_col = (int32_t)hot.y;
                } else {
                    lv = linear_half_view(job, row0, lane, m0 + quad * 32);
                    f32_tma = job.epi == EPI_LINEAR && !job.out_half && (job.out_map1 > 0 || job.out2_map1 > 0);
                    f32_scale = job.out_scale != 0.0f ? job.out_scale : 1.0f;
                }
                if (et == 0) trace_stamp(trace, 2, seq, 0);
                if (kCluster == 2) mbar_wait_guarded(&ctl->tmem_full[buf], (seq >> 1) & 1u);
                else mbar_wait(&ctl->tmem_full[buf], (seq >> 1) & 1u);
                tcgen05_fence_after();
                if (et == 0) trace_stamp(trace, 2, seq, 1);
                // (tried and dropped: software-pipelining the accumulator reads -- 16 columns at a time, or the next chunk's read
                //  issued before the current chunk is staged.  With the read's registers live across the loop ptxas feeds the
                //  eight bias loads one by one into the additions: 11 % slower on a [131072 x 512] . [512 x 512] layer.)
                if (fast_linear && !(debug_mode & (4 | 32768))) {
                    // Both 64-column reads of this warp's share, software-pipelined: the second one is in flight while the first
                    // block is turned into fp16 and stored.  Tensor memory delivers 64 B per cycle and SM; with all eight warps
                    // reading at the same moment (they are released together by tmem_full) and then all computing, the 2048
                    // cycles of reads and the ~1600 of arithmetic of a tile simply added up (timeline: 3685 cycles of work per job).
                    float va[64], vb[64];
                    auto store_block = [&](float (&v)[64], int c0) {
                        uint32_t pk[32];
                        linear_half_pack64_smem(lv.alpha, v, bias_sa, c0, pk);
                        const uint32_t tile = next_store_tile(my_stage);
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            sts128(tile + lane * 128 + ((q ^ (lane & 7)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                        fence_async_shared();
                        __syncwarp();
                        if (lane == 0 && !(debug_mode & 16)) {
                            if (out_scratch && !(debug_mode & 8192)) tma_store_2d_hint(&maps[lv.out_map], tile, lv.out_col + c0, row0, keep_policy);
                            else tma_store_2d(&maps[lv.out_map], tile, lv.out_col + c0, row0);
                            bulk_commit();
                        }
                        tma_pending = true;
                    };
                    tmem_load_64cols_nowait(taddr + (uint32_t)c_begin, va);
                    tmem_load_wait();
                    tmem_load_64cols_nowait(taddr + (uint32_t)(c_begin + 64), vb);
                    store_block(va, c_begin);
                    tmem_load_wait();
                    store_block(vb, c_begin + 64);
                } else if (lstm_fast) {
                    const int row = row0 + lane, unit0 = lstm_unit_of_packed(job.n_begin + c_begin);
                    const bool in_rows = row < job.m_rows;
                    const bool live = in_rows && job.t < lf_seq_len;
                    const uint32_t tile = smem_addr_of(my_stage);
                    mbar_wait(&ctl->c_bar[ew], c_phase);
                    c_phase ^= 1u;
                    uint32_t hp[16];                       // this row's 32 new hidden states, fp16
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {       // 64 gate columns = 16 units at a time
                        float v[64], bb[64];
                        tmem_load_64cols_nowait(taddr + (uint32_t)(c_begin + 64 * hf), v);
                        bias_from_smem<64>(bias_sa, c_begin + 64 * hf, bb);
                        float c_old[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t = lds128f(tile + lane * 128 + (((4 * hf + q) ^ (lane & 7)) << 4));
                            c_old[4 * q] = t.x; c_old[4 * q + 1] = t.y; c_old[4 * q + 2] = t.z; c_old[4 * q + 3] = t.w;
                        }
                        tmem_load_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) add_f32x2(v[2 * i], v[2 * i + 1], bb[2 * i], bb[2 * i + 1]);
                        float c_new[16], h[16];
#pragma unroll
                        for (int g = 0; g < 2; ++g)
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                lstm_cell_fast(v[32 * g + k], v[32 * g + 8 + k], v[32 * g + 16 + k], v[32 * g + 24 + k], c_old[8 * g + k],
                                               c_new[8 * g + k], h[8 * g + k]);
                        if (live) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                sts128f(tile + lane * 128 + (((4 * hf + q) ^ (lane & 7)) << 4),
                                        make_float4(c_new[4 * q], c_new[4 * q + 1], c_new[4 * q + 2], c_new[4 * q + 3]));
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __half2 hh = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
                            hp[8 * hf + i] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                    }
                    const bool h_tma = job.out_map1 > 0 && !(debug_mode & 65536);
                    if (in_rows && !live && !(debug_mode & 16)) {       // the sequence has ended: the state is carried (packed-sequence semantics); c stays as it is in the tile
                        const uint4* hprev = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(job.h_prev) + (int64_t)row * job.h_prev_stride + unit0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 t = __ldcg(hprev + q);
                            hp[4 * q] = t.x; hp[4 * q + 1] = t.y; hp[4 * q + 2] = t.z; hp[4 * q + 3] = t.w;
                        }
                    }
                    if (h_tma) {
                        // the new hidden states leave by TMA as well ([32 rows x 64 B] in the warp's second staging tile, 64-byte
                        // swizzle): nothing of this job is stored from registers, so its publication needs no proxy fence
                        const uint32_t htile = tile + (uint32_t)(kStageFloats * 4);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            sts128(htile + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
                    } else if (in_rows && !(debug_mode & 16)) {
                        uint4* hout = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(job.out) + (int64_t)row * job.out_stride + unit0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) hout[q] = make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
                    }
                    if (!h_tma) generic_stores = true;
                    fence_async_shared();
                    __syncwarp();
                    if (lane == 0 && !(debug_mode & 16)) {
                        tma_store_2d(&maps[job.c_map1 - 1], tile, unit0, row0);
                        if (h_tma) tma_store_2d(&maps[job.out_map1 - 1], tile + (uint32_t)(kStageFloats * 4), unit0, row0);
                        bulk_commit();
                    }
                    tma_pending = true;
                } else
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    if (kWideChunks && c0 + 64 <= lv.fast_cols && c0 + 64 <= c_end && !(debug_mode & 1024)) {
                        // fp16 linear jobs, 64 columns at a time: ONE accumulator read (one wait) per 64 columns, the bias from
                        // shared memory, packed fp32 arithmetic, and the [32 rows x 128 B] block leaves through ONE TMA store
                        // from the warp's staging tile (no shared-memory read-back, no per-lane global stores).
                        if (lv.out_map >= 0 && !(debug_mode & 2048)) {
                            float v2[64], bb[64];
                            tmem_load_64cols_nowait(taddr + (uint32_t)c0, v2);
                            bias_from_smem<64>(bias_sa, c0, bb);
                            tmem_load_wait();
                            if (!(debug_mode & 4)) {
                                uint32_t pk[32];
                                linear_half_pack64(lv, v2, bb, pk);
                                const uint32_t tile = next_store_tile(my_stage);
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    sts128(tile + lane * 128 + ((q ^ (lane & 7)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                                fence_async_shared();
                                __syncwarp();
                                if (lane == 0 && !(debug_mode & 16)) {
                                    if (out_scratch && !(debug_mode & 8192)) tma_store_2d_hint(&maps[lv.out_map], tile, lv.out_col + c0, row0, keep_policy);
                                    else tma_store_2d(&maps[lv.out_map], tile, lv.out_col + c0, row0);
                                    bulk_commit();
                                }
                                tma_pending = true;
                            }
                            c0 += 32;
                            continue;
                        }
                        release_tile();
                        float v2[64], b0[32], b1[32];
                        tmem_load_64cols_nowait(taddr + (uint32_t)c0, v2);
                        bias_from_smem<32>(bias_sa, c0, b0);
                        bias_from_smem<32>(bias_sa, c0 + 32, b1);
                        tmem_load_wait();
                        if (!(debug_mode & 4)) {
                            uint32_t pk[16];
                            linear_half_pack(lv, *reinterpret_cast<const float(*)[32]>(&v2[0]), b0, pk);
                            linear_half_store(lv, row0, lane, c0, pk, my_stage, (debug_mode & 16) != 0);
                            linear_half_pack(lv, *reinterpret_cast<const float(*)[32]>(&v2[32]), b1, pk);
                            linear_half_store(lv, row0, lane, c0 + 32, pk, my_stage, (debug_mode & 16) != 0);
                        }
                        generic_stores = true;
                        c0 += 32;
                        continue;
                    }
                    if (f32_tma && !(debug_mode & 2048)) {
                        // fp32 outputs of a plain contraction (the blend GEMMs): the [32 rows x 32 columns] block goes through the
                        // swizzled staging tile and ONE TMA store, which also clips rows and columns beyond the valid ones
                        const int n0 = job.n_begin + c0;
                        const bool second = n0 >= job.split;
                        const int map1 = second ? job.out2_map1 : job.out_map1;
                        if (map1 > 0 && (second || n0 + 32 <= job.split || job.split >= job.n_valid)) {
                            float v[32], bb[32];
                            tmem_load_32cols_nowait(taddr + (uint32_t)c0, v);
                            bias_from_smem<32>(bias_sa, c0, bb);
                            tmem_load_wait();
                            if (!(debug_mode & 4)) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    v[i] = fmaf(v[i], f32_scale, bb[i]);
                                    if (job.round_out) v[i] = round_tf32(v[i]);
                                }
                                const uint32_t tile = next_store_tile(my_stage);
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    sts128f(tile + lane * 128 + ((q ^ (lane & 7)) << 4), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                                fence_async_shared();
                                __syncwarp();
                                if (lane == 0 && !(debug_mode & 16)) {
                                    tma_store_2d(&maps[map1 - 1], tile, second ? n0 - job.split : job.out_col0 + n0, row0);
                                    bulk_commit();
                                }
                                tma_pending = true;
                            }
                            continue;
                        }
                    }
                    release_tile();
                    float v[32];
                    if (c0 + 32 <= lv.fast_cols) {
                        float bias[32];
                        bias_from_smem<32>(bias_sa, c0, bias);
                        tmem_load_32cols(taddr + (uint32_t)c0, v);
                        if (!(debug_mode & 4)) linear_half_chunk(lv, row0, lane, c0, v, bias, my_stage, (debug_mode & 16) != 0);
                        generic_stores = true;
                        continue;
                    }
                    if (debug_mode & 32) {                 // measurement only: no TMEM read
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                    } else {
                        tmem_load_32cols(taddr + (uint32_t)c0, v);
                    }
                    if (!(debug_mode & 4))
                        epilogue_chunk(job, row0, lane, c0, v, my_stage, (debug_mode & 16) != 0, lstm_pre ? cpre : nullptr, bias_sa, m0 + quad * 32);
                    generic_stores = true;
                    // ... and the next pair's while this pair is being computed
                    if (lstm_pre && !(c0 & 32) && c0 + 64 < c_end) lstm_half_load_c(job, row0, lane, c0 + 64, cpre);
                }
                // ---- the accumulator goes back to the MMA issuer as soon as this warp has read its lanes ----
                tcgen05_fence_before();
                __syncwarp();
                if (et == 0) trace_stamp(trace, 2, seq, 2);
                if (lane == 0) {
                    if (kPair) mbar_arrive_cluster(mapa_rank0(&ctl->tmem_empty[buf]));      // CTA 0 issues the MMAs of both
                    else mbar_arrive(&ctl->tmem_empty[buf]);
                }
                // ---- publication: only a job some later job waits for hands its stores over ----
                const int is_dep = fast_linear ? (int)((hot.x >> 9) & 3u) : job.is_dep;
                uint32_t* const done_ctr = fast_linear ? nullptr : job.done_ctr;
                if (!(debug_mode & 8) && (is_dep || done_ctr)) {
                    // Blocks that left through TMA stores are in the async proxy already: their issuing lane has to see the
                    // writes complete.  Stores from registers (any other path, this job's or an earlier one's: a thread owns
                    // the same rows and columns in all jobs of an item) are ordered at CTA scope and handed to the async
                    // proxy by the thread that made them.
                    const bool had_generic = generic_stores;      // warp-uniform: the paths above are
                    if (generic_stores) {
                        __threadfence_block();
                        asm volatile("fence.proxy.async;" ::: "memory");
                        generic_stores = false;
                    }
                    __syncwarp();
                    if (is_dep == 2 && !had_generic && !done_ctr) {
                        // the dependent job is at least two jobs away: its producer can wait for the completion to be noticed at
                        // this warp's next staging-tile wait (release_tile)
                        if (pub_pending && lane == 0) { bulk_wait_all0(); ctl->done_seq[ew] = pub_pending; }      // (an earlier one still open: close it now)
                        pub_pending = seq + 1u;
                    } else {
                        if (lane == 0) {
                            bulk_wait_all0();
                            if (is_dep) ctl->done_seq[ew] = seq + 1u;
                        }
                        pub_pending = 0;
                        tma_pending = false;
                        if (done_ctr) {
                            // cross-CTA consumers (the LSTM wavefront): the LAST epilogue warp to get here publishes the tile;
                            // the acq_rel count makes the other warps' stores part of what its release covers
                            if (lane == 0) {
                                uint32_t old;
                                asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&ctl->job_done_cnt[seq & 3u])) : "memory");
                                if ((old + 1u) % (uint32_t)kEpiWarps == 0u && m0 < m_tiles * kTileM) red_release_gpu(done_ctr + m0 / kTileM, 1u);
                            }
                        }
                    }
                } else if (pub_pending && c_begin >= c_end) {
                    release_tile();          // (a warp without columns in this job never reaches the staging-tile wait)
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->job_empty[buf]);
                if (et == 0) trace_stamp(trace, 2, seq, 3);
            }
        }
        release_tile();
        if (lane == 0) bulk_wait_all0();          // no TMA store may still read this CTA's shared memory when it exits
    }

    tcgen05_fence_before();
    __syncthreads();
    if (kCluster == 2) cluster_sync_all();        // nobody leaves while the peer may still multicast into this CTA
    if (warp == 2) {
        tcgen05_fence_after();
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

int tc_encode_map(void* out_map, const float* base, int64_t row_stride_elems, int k_extent, int64_t rows,
                  int box_rows, int half) {
    const int64_t row_stride_floats = row_stride_elems;      // (name kept for the messages below)
    const int elem = half ? 2 : 4;
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return EMPOSE_E_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((row_stride_floats * elem) & 15) || box_rows < 1 || box_rows > 256) {
        set_last_error("tensor map: base / stride must be 16-byte aligned and the box at most 256 rows");
        return EMPOSE_E_ARG;
    }
    cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_floats * elem};
    // half == 2: fp16 elements in boxes of 32 (64 bytes, 64-byte swizzle): the hidden-state block of an LSTM epilogue warp
    cuuint32_t box[2] = {(cuuint32_t)(half == 2 ? 32 : half ? kChunkKHalf : kChunkK), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    half == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return EMPOSE_E_CUDA;
    }
    return EMPOSE_OK;
}

// Executor configuration, once per process (one process per GPU).
static int g_max_clusters = 0;          // co-resident 2-CTA clusters (0: cluster modes unavailable or disabled)
static int g_cluster_mode = 0;          // 2: W multicast, 3: CTA-pair MMA
// EMPOSE_TC_DEBUG bits.  Throughput experiments (results are garbage): 1 TMA only, 2 MMA only, 4 no epilogue math, 8 no
// publication, 16 no stores, 32 no TMEM reads.  A/B switches (results stay correct): 128 no cell-state prefetch (and with it the
// chunk-wise LSTM epilogue), 256 CTA-pair mode with a 3-slot ring, 1024 32-column chunks only, 2048 stores from registers instead
// of TMA stores, 4096 chunk-wise LSTM epilogue, 8192 no L2 hints, 16384 no one-load job summary, 32768 accumulator reads not
// pipelined, 65536 LSTM hidden state stored from registers, 131072 no tensor-map prefetch, 262144 all-spaces proxy fences
static int g_debug_mode = 0;
static unsigned long long* g_trace_dev = nullptr;      // EMPOSE_TC_TRACE: the stamps of the LAST launch are written to that file
static const char* g_trace_path = nullptr;
constexpr size_t kTraceWords = (size_t)kTraceCtas * kTraceRoles * kTraceJobs * kTraceStamps;

static int g_trace_kind = 0;            // EMPOSE_TC_TRACE_KIND: 0 any launch, 1 item-table launches (LSTM wavefront), 2 chained items (MLP chains)
static bool tc_trace_wanted(int kind) { return g_trace_dev && (g_trace_kind == 0 || g_trace_kind == kind); }
static void tc_trace_begin(cudaStream_t s, int kind) {
    if (tc_trace_wanted(kind)) cudaMemsetAsync(g_trace_dev, 0, kTraceWords * 8, s);
}
static void tc_trace_end(cudaStream_t s, int kind) {
    if (!tc_trace_wanted(kind)) return;
    std::vector<unsigned long long> h(kTraceWords);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), g_trace_dev, kTraceWords * 8, cudaMemcpyDeviceToHost);
    FILE* f = fopen(g_trace_path, "w");
    if (!f) return;
    for (int c = 0; c < kTraceCtas; ++c)
        for (int r = 0; r < kTraceRoles; ++r)
            for (int j = 0; j < kTraceJobs; ++j) {
                const unsigned long long* t = &h[(((size_t)c * kTraceRoles + r) * kTraceJobs + j) * kTraceStamps];
                if (t[0] | t[1] | t[2] | t[3]) fprintf(f, "%d %d %d %llu %llu %llu %llu\n", c, r, j, t[0], t[1], t[2], t[3]);
            }
    fclose(f);
}

static int tc_configure(int num_sms) {
    static bool configured = false;
    if (configured) return EMPOSE_OK;
    EMPOSE_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    EMPOSE_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    EMPOSE_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const char* e = getenv("EMPOSE_TC_DEBUG");
    g_debug_mode = e ? atoi(e) : 0;
    // EMPOSE_TC_CLUSTER: 0 one CTA per row tile; 1 pairs of CTAs with the W tile multicast (correct, halves the W
    // traffic out of L2, no faster: profiles/r01/README.md); 2 pairs of CTAs issuing ONE cta_group::2 MMA per pair
    // (each SM stages half of W: relieves the shared-memory bandwidth that bounds the mainloop).
    const char* c = getenv("EMPOSE_TC_CLUSTER");
    const int want = c ? atoi(c) : kDefaultClusterMode;
    if ((want == 1 || want == 2) && !(g_debug_mode & 3)) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (num_sms / 2)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int n = 0;
        const cudaError_t q = want == 1 ? cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<2>, &cfg)
                                        : cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<3>, &cfg);
        if (q == cudaSuccess && n > 0) { g_max_clusters = n; g_cluster_mode = want + 1; }
        else cudaGetLastError();
    }
    g_trace_path = getenv("EMPOSE_TC_TRACE");
    if (g_trace_path && *g_trace_path) {
        EMPOSE_CUDA_TRY(cudaMalloc(&g_trace_dev, kTraceWords * 8));
        EMPOSE_CUDA_TRY(cudaMemcpyToSymbol(g_trace_buf, &g_trace_dev, sizeof(g_trace_dev)));
        const char* k = getenv("EMPOSE_TC_TRACE_KIND");
        g_trace_kind = k ? atoi(k) : 0;
    }
    if (getenv("EMPOSE_TC_VERBOSE"))
        fprintf(stderr, "empose_b200: gemm executor: mode %d, %d co-resident 2-CTA clusters on %d SMs\n", g_cluster_mode, g_max_clusters, num_sms);
    configured = true;
    return EMPOSE_OK;
}
#define EMPOSE_TRY_CONFIG(num_sms)                 \
    do {                                           \
        int _rc = tc_configure(num_sms);           \
        if (_rc != EMPOSE_OK) return _rc;          \
    } while (0)

static int tc_launch_inner(const GemmJob* d_jobs, const void* d_maps, int job_begin, int job_count, int jobs_per_item, int m_tiles,
                           int num_sms, cudaStream_t stream) {
    const int max_clusters = g_max_clusters, cluster_mode = g_cluster_mode, debug_mode = g_debug_mode;
    const int2* no_items = nullptr;
    const int groups = job_count / jobs_per_item;
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(d_maps);
    if (max_clusters > 0 && m_tiles >= 2) {
        const int n_items = ((m_tiles + 1) / 2) * groups;
        const int clusters = n_items < max_clusters ? n_items : max_clusters;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (cluster_mode == 3)
            EMPOSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<3>, d_jobs, maps, job_begin, job_count, jobs_per_item, m_tiles, debug_mode, no_items, 0, 0u));
        else
            EMPOSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2>, d_jobs, maps, job_begin, job_count, jobs_per_item, m_tiles, debug_mode, no_items, 0, 0u));
        return EMPOSE_OK;
    }
    const int n_items = m_tiles * groups;
    const int grid = n_items < num_sms ? n_items : num_sms;
    gemm_tc_kernel<1><<<grid, kThreads, kSmemBytes, stream>>>(d_jobs, maps, job_begin, job_count, jobs_per_item, m_tiles, debug_mode, no_items, 0, 0u);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int tc_item_rows(int m_tiles, int num_sms) {
    EMPOSE_TRY_CONFIG(num_sms);
    return (g_max_clusters > 0 && m_tiles >= 2) ? 2 * kTileM : kTileM;
}

static int tc_launch_items_inner(const GemmJob* d_jobs, const void* d_maps, const void* d_items, int n_items, int rows_per_unit, uint32_t epoch,
                                 int m_tiles, int num_sms, cudaStream_t stream) {
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(d_maps);
    const int2* items = reinterpret_cast<const int2*>(d_items);
    if (rows_per_unit != tc_item_rows(m_tiles, num_sms)) {
        set_last_error("internal error: item table built for a different executor mode");
        return EMPOSE_E_ARG;
    }
    // The grid must be co-resident: CTAs spin on counters that other CTAs of the same launch advance.  Items are taken
    // round-robin in table order and every item only waits for items earlier in the table, so the earliest unfinished item
    // can always run.
    if (rows_per_unit == 2 * kTileM) {
        const int clusters = n_items < g_max_clusters ? n_items : g_max_clusters;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (g_cluster_mode == 3)
            EMPOSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<3>, d_jobs, maps, 0, n_items, 1, m_tiles, g_debug_mode, items, n_items, epoch));
        else
            EMPOSE_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2>, d_jobs, maps, 0, n_items, 1, m_tiles, g_debug_mode, items, n_items, epoch));
        return EMPOSE_OK;
    }
    const int grid = n_items < num_sms ? n_items : num_sms;
    gemm_tc_kernel<1><<<grid, kThreads, kSmemBytes, stream>>>(d_jobs, maps, 0, n_items, 1, m_tiles, g_debug_mode, items, n_items, epoch);
    EMPOSE_CUDA_TRY(cudaGetLastError());
    return EMPOSE_OK;
}

int tc_launch(const GemmJob* d_jobs, const void* d_maps, int job_begin, int job_count, int jobs_per_item, int m_tiles,
              int num_sms, cudaStream_t stream) {
    EMPOSE_TRY_CONFIG(num_sms);
    const int kind = jobs_per_item > 1 ? 2 : 3;
    tc_trace_begin(stream, kind);
    const int rc = tc_launch_inner(d_jobs, d_maps, job_begin, job_count, jobs_per_item, m_tiles, num_sms, stream);
    tc_trace_end(stream, kind);
    return rc;
}

int tc_launch_items(const GemmJob* d_jobs, const void* d_maps, const void* d_items, int n_items, int rows_per_unit, uint32_t epoch,
                    int m_tiles, int num_sms, cudaStream_t stream) {
    EMPOSE_TRY_CONFIG(num_sms);
    tc_trace_begin(stream, 1);
    const int rc = tc_launch_items_inner(d_jobs, d_maps, d_items, n_items, rows_per_unit, epoch, m_tiles, num_sms, stream);
    tc_trace_end(stream, 1);
    return rc;
}

}  // namespace empose
