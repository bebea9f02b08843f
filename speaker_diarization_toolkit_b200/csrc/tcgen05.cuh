// tcgen05.cuh -- PTX wrappers (mbarrier, TMA, tcgen05.mma/ld/commit, UMMA descriptors), the kernel parameter block and
// the candidate flush shared by the tcgen05 kernels (poolgemm.cu: columns = segments; poolacc.cu: columns = label groups).
#pragma once
#include <cuda.h>

#include "common.cuh"

#define PG_THREADS 192    // MT == 1: TMA warp, MMA warp, one epilogue warpgroup
#define PG_THREADS2 320   // MT == 2: two epilogue warpgroups, one per row tile
#define PG_CS 16            // candidate slots per (label group, row block, row tile, warp) = per 32 bank rows
#define PG_SMEM_LIMIT 232448

struct PgParams {
    const int64_t* goff;
    const int32_t* range_g;     // [n_ranges+1] first group of each column range
    int32_t n_ranges, RB;
    int64_t P;
    int32_t g_base;             // first group of this batch (slot arrays are batch-relative)
    int32_t pool;
    float tau;
    int32_t mode;               // 0 candidates, 1 dense
    int32_t* slot_cnt;
    int32_t* slot_row;
    float* slot_val;
    float* slot_bound;
    float* dense_out;
    int32_t dense_ld;
    // Running per-label lower bound on the 64th best stage-A score (un-thresholded / low-threshold queries): 64 buckets
    // per slot group, bucket b = max key flushed so far by the sub-slots with index % 64 == b.  The 64 bucket maxima are
    // 64 different bank rows, so their minimum can only be <= the 64th best score of the label: rows below it can never
    // be among the (at most 64) candidates and are not written at all.  null = off.
    uint32_t* kth;
    int64_t nsub;               // sub-slots per slot group
};
#define PG_KTH 64

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pg_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void pg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pg_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pg_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PG_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PG_DONE_%=;\n\t"
        "bra PG_WAIT_%=;\n\t"
        "PG_DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void pg_tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void pg_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void pg_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// One lane of a CONVERGED warp.  The MMA-issuing warp runs its loop with all 32 lanes (loop counters, barrier addresses
// and UMMA descriptors are then warp-uniform and live in uniform registers) and only the tcgen05.mma / tcgen05.commit
// instructions are issued by the elected lane; a loop entered by lane 0 alone made the compiler re-elect and broadcast
// every descriptor (ELECT + R2UR.BROADCAST per MMA), ~700 cycles of issue latency per 512 tensor cycles at D = 512.
__device__ __forceinline__ bool pg_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void pg_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void pg_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void pg_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void pg_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cta_group::2 (a cluster of two CTAs on one TPC issues 256-row MMAs; poolgemm.cu k_poolgemm2, poolacc.cu k_poolacc2) ----
__device__ __forceinline__ uint32_t pg_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void pg_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pg_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void pg_mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void pg_tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tmap, int32_t c0, int32_t c1, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ void pg_mma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void pg_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (see cute/arch/mma_sm100_desc.hpp):
// start>>4 | LBO(=1, unused for swizzled K-major)<<16 | SBO(=1024B: 8 rows x 128B)>>4<<32 | version 1<<46 | SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t pg_make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// ---- the flush of one label group for this warp's 32 rows of row tile rt --------------------------
// sub-slot = (label group, row block, row tile, warp): at most PG_CS rows with approx >= tau are kept; if
// more pass, the PG_CS largest (rank of the orderable key among the warp's 32) and the bound below which rows
// were dropped.
// writes the candidates of one sub-slot; `mpass` = ballot of the lanes whose value passes (must be non-zero)
__device__ __forceinline__ void pg_flush_write(const PgParams& p, float val, bool pass, uint32_t mpass, int64_t sub, int32_t lane, int64_t row) {
    const int npass = __popc(mpass);
    if (lane == 0) p.slot_cnt[sub] = npass;
    const uint32_t lt = (1u << lane) - 1u;
    if (npass <= PG_CS) {
        if (pass) {
            const int pos = __popc(mpass & lt);
            p.slot_row[sub * PG_CS + pos] = (int32_t)row;
            p.slot_val[sub * PG_CS + pos] = val;
        }
        return;
    }
    // more than PG_CS rows pass: keep the PG_CS largest.  Every lane ranks its key among the 32 (32 independent shuffles,
    // ties by lane) -- a throughput-bound ~130 instructions instead of a 32-step dependent ballot search; the rank is
    // also the slot position, so the slot comes out sorted by descending score.
    const uint32_t key = pass ? sdk_fkey(val) : 0u;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint32_t kj = __shfl_sync(0xffffffffu, key, j);
        rank += (kj > key || (kj == key && j < lane)) ? 1 : 0;
    }
    const uint32_t mlast = __ballot_sync(0xffffffffu, rank == PG_CS - 1);       // exactly one lane: ranks are a permutation
    const uint32_t T = __shfl_sync(0xffffffffu, key, __ffs(mlast) - 1);         // smallest kept key: dropped rows are <= it
    if (lane == 0) p.slot_bound[sub] = sdk_funkey(T);
    if (rank < PG_CS) {
        p.slot_row[sub * PG_CS + rank] = (int32_t)row;
        p.slot_val[sub * PG_CS + rank] = val;
    }
}
// out-of-line copy for call sites that are unrolled many times (poolacc.cu: one per accumulator column)
static __device__ __noinline__ void pg_flush_write_call(const PgParams* p, float val, bool pass, uint32_t mpass, int64_t sub, int32_t lane, int64_t row) {
    pg_flush_write(*p, val, pass, mpass, sub, lane, row);
}

// running k-th best: threshold key of slot group sg (0 = nothing known yet)
__device__ __forceinline__ uint32_t pg_kth_load(const PgParams& p, int64_t sg, int32_t lane) {
    const uint32_t* kb = p.kth + sg * PG_KTH;
    const uint32_t a = __ldcg(kb + lane), b = __ldcg(kb + lane + 32);     // L2: other SMs raise the buckets with atomics
    return __reduce_min_sync(0xffffffffu, a < b ? a : b);
}
__device__ __forceinline__ void pg_kth_update(const PgParams& p, int64_t sg, int64_t si, uint32_t key_if_live, uint32_t thr, int32_t lane) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, key_if_live);
    if (lane == 0 && wmax > thr) atomicMax(p.kth + sg * PG_KTH + (si & (PG_KTH - 1)), wmax);
}

// sg = slot group (label group or accumulator column, batch-relative), si = sub-slot inside the group
__device__ __forceinline__ void pg_flush(const PgParams& p, float accv, int32_t g, int64_t n_g, int64_t sg, int64_t si, int32_t lane,
                                         int64_t row) {
    const float val = p.pool == 0 ? accv * (1.0f / (float)n_g) : accv;
    if (p.mode == 1) {
        if (row < p.P) p.dense_out[row * (int64_t)p.dense_ld + g] = val;
        return;
    }
    const int64_t sub = sg * p.nsub + si;
    bool pass = (val >= p.tau) && (row < p.P);
    if (p.kth) {
        const uint32_t thr = pg_kth_load(p, sg, lane);
        const uint32_t key = sdk_fkey(val);
        pg_kth_update(p, sg, si, pass ? key : 0u, thr, lane);
        pass = pass && key >= thr;
    }
    const uint32_t mpass = __ballot_sync(0xffffffffu, pass);
    if (lane == 0 && mpass == 0) p.slot_cnt[sub] = 0;
    if (mpass == 0) return;
    pg_flush_write(p, val, pass, mpass, sub, lane, row);
}

__device__ __forceinline__ void pg_tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}



// ---- host: TMA tensor map of a [rows, Dp] bf16 row-major matrix, box = 64 elements x box_rows, 128B swizzle ----
typedef CUresult (*pg_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int pg_make_tmap(sdk_ctx* c, CUtensorMap* tm, const void* base, int64_t rows, int32_t Dp, uint32_t box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)Dp, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)Dp * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((pg_encode_fn)c->tmap_encode)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sdk_fail(c, SDK_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return SDK_OK;
}


// merge of the per-(32 bank rows) candidate slots of one label group (poolgemm.cu); `sorted_group` (may be null) maps
// the launch index to the label group the result is written for
void pg_launch_merge(sdk_ctx* c, const int64_t* d_goff, int32_t g_base, int32_t ngroups, int32_t nsub, const int32_t* d_sorted_group,
                     float tau, int32_t ncand, int32_t* d_cand_row, float* d_gbound, const int32_t* d_glist = nullptr,
                     const int32_t* d_group_col = nullptr);
