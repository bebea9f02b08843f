// poolgemm.cu -- K2b + K3(stage A): segment x profile cosine GEMM on tcgen05 tensor cores with the
// per-label pooling fused into the epilogue.  The N x P score matrix never leaves TMEM/registers.
//
// Orientation.  The accumulator tile is  S^T = Bank(128 rows = TMEM lanes) x Segments(NC columns):
//   A operand (M side)  = bank rows,    bf16, K-major, resident in shared memory for a whole work unit
//   B operand (N side)  = segments,     bf16, K-major, streamed through a TMA ring, one 64-wide K chunk per stage
// so that pooling over a label's segments is a reduction ALONG TMEM COLUMNS: after tcgen05.ld each epilogue
// thread owns one bank row and simply adds (or maxes) the columns it receives into a running register --
// no cross-lane traffic.  Segments are sorted by label group, a work unit's column range is made of whole
// groups, and the running accumulator is flushed when the column index crosses a group end (goff[g+1]).
//
// Work unit = (column range of whole label groups, row block of MT*128 bank rows).  Units of the same column
// range are adjacent in the schedule, so the CTAs streaming the same segments run together and share them in
// L2; HBM sees every segment about once.  L2->SM operand traffic is 2/(MT*128) bytes per MAC.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
// (warp w owns TMEM lanes 32*(w%4)..+31).  Pipelines: bank tile full/empty, B ring full/empty, and a two-deep
// TMEM accumulator ring (2 x MT x NC columns <= 512) so the epilogue of chunk j overlaps the MMAs of chunk j+1.
//
// Flush (candidate mode): the warp keeps, for its 32*MT rows of label group g, the rows whose approximate
// pooled score is >= tau (at most CS of them; if more pass, the CS largest by rank among the warp's 32 keys, plus
// the bound below which everything was dropped).  A merge kernel then
// picks the top `ncand` rows per label and the upper bound on every row left out; select.cu re-scores the
// candidates canonically and checks the bound (the top-k certificate).
// Flush (dense mode, config 5): out[row, g] = pooled value.
#include <stdlib.h>

#include "tcgen05.cuh"

// ---- epilogue of one job: pool NC accumulator columns of this thread's bank row -------------------
// Pooling state of one row tile: current label group, its column range, the prefetched end of the NEXT
// group (so the goff load is off the critical path at a group boundary) and the running sum / max.
struct PgPool {
    int32_t g;
    int64_t gbeg, gend, gend_next;
    float acc;
};

__device__ __forceinline__ void pg_pool_init(PgPool& st, const PgParams& p, int32_t g_lo, int32_t g_hi, int64_t c0) {
    int32_t g0 = g_lo;
    int64_t e0 = p.goff[g0 + 1];
    while (e0 == c0 && g0 + 1 < g_hi) { ++g0; e0 = p.goff[g0 + 1]; }   // skip empty groups
    st.g = g0;
    st.gbeg = c0;
    st.gend = e0;
    st.gend_next = (g0 + 1 < g_hi) ? p.goff[g0 + 2] : 0x7fffffffffffffffLL;
    st.acc = p.pool == 0 ? 0.f : -3.0e38f;
}

// Reductions of one 32-column block as balanced trees (depth 5): one epilogue warp sits alone on its scheduler, so a
// serial 32-long fmax / fadd chain is pure latency (4 cycles a link, ~130 cycles per block -- it made max pooling
// epilogue-bound at D <= 256); the trees expose 16 independent operations per level instead.
__device__ __forceinline__ float pg_tree_sum(const float (&w)[32]) {
    float s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = (w[4 * q] + w[4 * q + 1]) + (w[4 * q + 2] + w[4 * q + 3]);
    return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
__device__ __forceinline__ float pg_tree_max(const float (&w)[32]) {
    float s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = fmaxf(fmaxf(w[4 * q], w[4 * q + 1]), fmaxf(w[4 * q + 2], w[4 * q + 3]));
    return fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), fmaxf(fmaxf(s[4], s[5]), fmaxf(s[6], s[7])));
}

__device__ __forceinline__ void pg_pool_block(PgPool& st, const PgParams& p, const float (&v)[32], int64_t cbase, int64_t c1,
                                              int32_t g_hi, int64_t sub_stride_base, int64_t subbase, int32_t lane, int64_t row) {
    const int64_t left = c1 - cbase;
    const int nvalid = left < 32 ? (int)left : 32;
    if (nvalid == 32 && st.gend > cbase + 32) {
        // whole block inside the current label group
        if (p.pool == 0) st.acc += pg_tree_sum(v);
        else st.acc = fmaxf(st.acc, pg_tree_max(v));
        return;
    }
    int c = 0;
    while (c < nvalid) {
        const int64_t togo = st.gend - cbase;
        const int run_end = togo < nvalid ? (int)togo : nvalid;
        // columns [c, run_end) of the block belong to the current group: mask, then the same trees
        float w[32];
        const unsigned span = (unsigned)(run_end - c);
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) {
            const bool on = (unsigned)(cc - c) < span;
            w[cc] = on ? v[cc] : (p.pool == 0 ? 0.f : -3.0e38f);
        }
        if (p.pool == 0) st.acc += pg_tree_sum(w);
        else st.acc = fmaxf(st.acc, pg_tree_max(w));
        c = run_end;
        if (cbase + run_end == st.gend) {
            pg_flush(p, st.acc, st.g, st.gend - st.gbeg, (int64_t)(st.g - p.g_base), subbase, lane, row);
            st.acc = p.pool == 0 ? 0.f : -3.0e38f;
            st.gbeg = st.gend;
            if (st.g + 1 < g_hi) {
                ++st.g;
                st.gend = st.gend_next;
                while (st.gend == st.gbeg && st.g + 1 < g_hi) { ++st.g; st.gend = p.goff[st.g + 1]; }
                st.gend_next = (st.g + 1 < g_hi) ? p.goff[st.g + 2] : 0x7fffffffffffffffLL;   // prefetch
            } else {
                st.gend = 0x7fffffffffffffffLL;   // past the last group of the unit
            }
        }
    }
}

// TMEM loads are software pipelined: the load of block b+1 is in flight while block b is pooled.
template <int NC>
__device__ __forceinline__ void pg_epi_job(PgPool& st, const PgParams& p, uint32_t taddr, int64_t cjob, int64_t c1, int32_t g_hi,
                                           int64_t sub_stride_base, int64_t subbase, int32_t lane, int64_t row) {
    float va[32], vb[32];
    pg_tmem_ld32(taddr, va);
#pragma unroll 1
    for (int blk = 0; blk < NC / 32; blk += 2) {
        const int64_t cb0 = cjob + blk * 32, cb1 = cb0 + 32;
        pg_tmem_ld_wait();
        if (cb1 < c1) pg_tmem_ld32(taddr + (blk + 1) * 32, vb);
        pg_pool_block(st, p, va, cb0, c1, g_hi, sub_stride_base, subbase, lane, row);
        if (cb1 >= c1) break;
        pg_tmem_ld_wait();
        if (blk + 2 < NC / 32 && cb1 + 32 < c1) pg_tmem_ld32(taddr + (blk + 2) * 32, va);
        pg_pool_block(st, p, vb, cb1, c1, g_hi, sub_stride_base, subbase, lane, row);
        if (cb1 + 32 >= c1) break;
    }
}

// ---- the kernel --------------------------------------------------------------------------------
// A "job" is (column chunk j, row tile rt): KCH*4 MMAs of 128 x NC x 16 into accumulator slot (job % NSLOT).
// Jobs run chunk-major, row-tile-minor, so with MT > 1 the whole B chunk (KCH stages) stays resident until its
// last row tile has consumed it (STAGES >= KCH + 1), and the epilogue of job i overlaps the MMAs of job i+1.
// N = NC = 256 keeps the per-MMA shared-memory operand fetch (4 KB of A + NC*32 B of B) under the MMA time.
template <int KCH, int MT, int NC, int STAGES>
__global__ void __launch_bounds__(MT == 2 ? PG_THREADS2 : PG_THREADS, 1)
k_poolgemm(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const PgParams p) {
    constexpr uint32_t A_TILE = 128 * 128;               // 128 rows x 64 bf16
    constexpr uint32_t A_BYTES = MT * KCH * A_TILE;
    constexpr uint32_t B_STAGE = NC * 128;
    constexpr uint32_t NSLOT = 512 / NC;                 // accumulator slots in TMEM
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    static_assert(NC == 128 || NC == 256, "accumulator slot width");
    static_assert(MT == 1 || STAGES >= KCH + 1, "B chunk must stay resident across the row tiles");
    static_assert(MT == 1 || (MT == 2 && NSLOT % 2 == 0), "two epilogue warpgroups: every accumulator slot belongs to one row tile");

    extern __shared__ uint8_t pg_smem_raw[];
    const uint32_t raw = pg_smem_u32(pg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + A_BYTES;
    const uint32_t sBar = sB + STAGES * B_STAGE;
    const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8;
    const uint32_t bar_b_full = sBar + 16, bar_b_empty = bar_b_full + 8 * STAGES;
    const uint32_t bar_t_full = bar_b_empty + 8 * STAGES, bar_t_empty = bar_t_full + 8 * NSLOT;
    const uint32_t s_tmem = bar_t_empty + 8 * NSLOT;
    uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(pg_smem_raw + (s_tmem - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        pg_mbar_init(bar_a_full, 1);
        pg_mbar_init(bar_a_empty, 1);
        for (int s = 0; s < STAGES; ++s) { pg_mbar_init(bar_b_full + 8 * s, 1); pg_mbar_init(bar_b_empty + 8 * s, 1); }
        for (uint32_t b = 0; b < NSLOT; ++b) { pg_mbar_init(bar_t_full + 8 * b, 1); pg_mbar_init(bar_t_empty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pg_fence_before();
    __syncthreads();
    pg_fence_after();
    const uint32_t tmem_base = *s_tmem_ptr;

    const int64_t n_units = (int64_t)p.n_ranges * p.RB;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_phase = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int32_t range = (int32_t)(u / p.RB), rb = (int32_t)(u - (int64_t)range * p.RB);
                const int64_t c0 = p.goff[p.range_g[range]], c1 = p.goff[p.range_g[range + 1]];
                if (c1 <= c0) continue;
                pg_mbar_wait(bar_a_empty, a_phase ^ 1);
                pg_mbar_expect_tx(bar_a_full, A_BYTES);
#pragma unroll 1
                for (int rt = 0; rt < MT; ++rt)
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc)
                        pg_tma_load_2d(sA + (rt * KCH + kc) * A_TILE, &tmapA, kc * 64, (int32_t)((int64_t)rb * MT * 128 + rt * 128), bar_a_full);
                a_phase ^= 1;
                const int64_t nchunks = (c1 - c0 + NC - 1) / NC;
                for (int64_t j = 0; j < nchunks; ++j) {
                    const int32_t crow = (int32_t)(c0 + j * NC);
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc) {
                        pg_mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
                        pg_mbar_expect_tx(bar_b_full + 8 * stage, B_STAGE);
                        pg_tma_load_2d(sB + stage * B_STAGE, &tmapB, kc * 64, crow, bar_b_full + 8 * stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop, one elected lane issues) =================
        {
            uint32_t stage = 0, phase = 0, a_phase = 0;
            uint32_t job = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int32_t range = (int32_t)(u / p.RB);
                const int64_t c0 = p.goff[p.range_g[range]], c1 = p.goff[p.range_g[range + 1]];
                if (c1 <= c0) continue;
                pg_mbar_wait(bar_a_full, a_phase);
                a_phase ^= 1;
                pg_fence_after();
                const int64_t nchunks = (c1 - c0 + NC - 1) / NC;
                for (int64_t j = 0; j < nchunks; ++j) {
                    const uint32_t stage0 = stage, phase0 = phase;       // ring position of this chunk's first K stage
#pragma unroll
                    for (int rt = 0; rt < MT; ++rt, ++job) {
                        const uint32_t slot = job % NSLOT, it = job / NSLOT;
                        pg_mbar_wait(bar_t_empty + 8 * slot, (it & 1u) ^ 1u);
                        pg_fence_after();
                        const uint32_t td = tmem_base + slot * NC;
                        uint32_t st = stage0, ph = phase0;
#pragma unroll 1
                        for (int kc = 0; kc < KCH; ++kc) {
                            if (rt == 0) {
                                pg_mbar_wait(bar_b_full + 8 * st, ph);
                                pg_fence_after();
                            }
                            const uint64_t db = pg_make_desc(sB + st * B_STAGE);
                            const uint64_t da = pg_make_desc(sA + (rt * KCH + kc) * A_TILE);
                            if (pg_elect_one()) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    pg_mma_bf16(td, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC, (kc | kk) != 0 ? 1u : 0u);
                                if (rt == MT - 1) pg_commit(bar_b_empty + 8 * st);   // last row tile: this K stage is free
                            }
                            __syncwarp();
                            if (++st == STAGES) { st = 0; ph ^= 1; }
                        }
                        if (pg_elect_one()) pg_commit(bar_t_full + 8 * slot);         // accumulator of this job complete
                        __syncwarp();
                        if (rt == MT - 1) { stage = st; phase = ph; }
                    }
                }
                if (pg_elect_one()) pg_commit(bar_a_empty);                           // bank tiles may be overwritten
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue: warpgroup wg handles row tile rt == wg (MT == 2) or all jobs (MT == 1);
        //                    thread <-> bank row (TMEM lane) =================
        const int wq = warp & 3;                                 // TMEM lane quadrant of this warp
        const int wg = (warp - 2) >> 2;                          // epilogue warpgroup 0/1
        const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
        uint32_t jobbase = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int32_t range = (int32_t)(u / p.RB), rb = (int32_t)(u - (int64_t)range * p.RB);
            const int32_t g_lo = p.range_g[range], g_hi = p.range_g[range + 1];
            const int64_t c0 = p.goff[g_lo], c1 = p.goff[g_hi];
            if (c1 <= c0) continue;
            const int rt = MT == 2 ? wg : 0;
            const int64_t tile128 = (int64_t)rb * MT + rt;
            const int64_t row = tile128 * 128 + wq * 32 + lane;
            const int64_t subbase = tile128 * 4 + wq;
            PgPool st;
            pg_pool_init(st, p, g_lo, g_hi, c0);
            const int64_t nchunks = (c1 - c0 + NC - 1) / NC;
            for (int64_t j = 0; j < nchunks; ++j) {
                const uint32_t job = jobbase + (uint32_t)j * MT + rt;
                const uint32_t slot = job % NSLOT, it = job / NSLOT;
                pg_mbar_wait(bar_t_full + 8 * slot, it & 1u);
                pg_fence_after();
                pg_epi_job<NC>(st, p, tmem_base + lane_base + slot * NC, c0 + j * NC, c1, g_hi, (int64_t)p.RB * MT * 4, subbase, lane, row);
                pg_fence_before();
                __syncwarp();
                if (lane == 0) pg_mbar_arrive(bar_t_empty + 8 * slot);
            }
            jobbase += (uint32_t)nchunks * MT;
        }
    }

    // ---- teardown ----
    pg_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// =================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC issues 256 x NC x 16 MMAs.  CTA r owns bank rows
// [.., r*128 .. r*128+127] of every 256-row tile (its half of A and of the accumulator) and HALF of every
// segment chunk (B is split along N across the pair), so per CTA the shared-memory operand fetch and the
// L2 -> SM traffic of the streamed operand are halved, and a B stage is NC/2 x 128 B: the ring is twice as deep.
// The leader CTA (rank 0) issues all MMAs; full barriers live in the leader and collect the TMA bytes of
// both CTAs; empty / accumulator-full barriers are signalled in both CTAs by multicast tcgen05.commit;
// the accumulator-empty barrier lives in the leader and is armed by the epilogue warps of both CTAs.
// =================================================================================================
template <int KCH, int MT, int NC, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MT == 2 ? PG_THREADS2 : PG_THREADS, 1)
k_poolgemm2(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const PgParams p) {
    constexpr uint32_t A_TILE = 128 * 128;               // this CTA's 128 rows x 64 bf16 of a 256-row tile
    constexpr uint32_t A_BYTES = MT * KCH * A_TILE;
    constexpr uint32_t B_HALF = (NC / 2) * 128;          // this CTA's half of a chunk's K stage
    constexpr uint32_t NSLOT = 512 / NC;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    static_assert(NC == 128 || NC == 256, "accumulator slot width");
    static_assert(MT == 1 || STAGES >= KCH + 1, "B chunk must stay resident across the row tiles");
    static_assert(MT == 1 || (MT == 2 && NSLOT % 2 == 0), "two epilogue warpgroups: every accumulator slot belongs to one row tile");

    extern __shared__ uint8_t pg_smem_raw[];
    const uint32_t raw = pg_smem_u32(pg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + A_BYTES;
    const uint32_t sBar = sB + STAGES * B_HALF;
    const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8;
    const uint32_t bar_b_full = sBar + 16, bar_b_empty = bar_b_full + 8 * STAGES;
    const uint32_t bar_t_full = bar_b_empty + 8 * STAGES, bar_t_empty = bar_t_full + 8 * NSLOT;
    const uint32_t s_tmem = bar_t_empty + 8 * NSLOT;
    uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(pg_smem_raw + (s_tmem - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = pg_cluster_rank();
    const bool leader = crank == 0;

    if (warp == 0 && lane == 0) {
        pg_mbar_init(bar_a_full, 1);
        pg_mbar_init(bar_a_empty, 1);
        for (int s = 0; s < STAGES; ++s) { pg_mbar_init(bar_b_full + 8 * s, 1); pg_mbar_init(bar_b_empty + 8 * s, 1); }
        for (uint32_t b = 0; b < NSLOT; ++b) { pg_mbar_init(bar_t_full + 8 * b, 1); pg_mbar_init(bar_t_empty + 8 * b, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    pg_fence_before();
    __syncthreads();
    pg_cluster_sync();                                   // peer barriers are initialised before anyone signals them
    pg_fence_after();
    const uint32_t tmem_base = *s_tmem_ptr;

    const int64_t n_units = (int64_t)p.n_ranges * p.RB;  // RB = row blocks of MT*256 rows
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    // leader-side addresses of the barriers that collect both CTAs' traffic
    const uint32_t l_a_full = pg_mapa(bar_a_full, 0), l_b_full = pg_mapa(bar_b_full, 0), l_t_empty = pg_mapa(bar_t_empty, 0);

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_phase = 0;
            for (int64_t u = pair; u < n_units; u += npairs) {
                const int32_t range = (int32_t)(u / p.RB), rb = (int32_t)(u - (int64_t)range * p.RB);
                const int64_t c0 = p.goff[p.range_g[range]], c1 = p.goff[p.range_g[range + 1]];
                if (c1 <= c0) continue;
                pg_mbar_wait(bar_a_empty, a_phase ^ 1);
                if (leader) pg_mbar_expect_tx(bar_a_full, 2 * A_BYTES);
#pragma unroll 1
                for (int rt = 0; rt < MT; ++rt)
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc)
                        pg_tma_load_2d_2sm(sA + (rt * KCH + kc) * A_TILE, &tmapA, kc * 64,
                                           (int32_t)(((int64_t)rb * MT + rt) * 256 + crank * 128), l_a_full);
                a_phase ^= 1;
                const int64_t nchunks = (c1 - c0 + NC - 1) / NC;
                for (int64_t j = 0; j < nchunks; ++j) {
                    const int32_t crow = (int32_t)(c0 + j * NC + crank * (NC / 2));
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc) {
                        pg_mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
                        if (leader) pg_mbar_expect_tx(bar_b_full + 8 * stage, 2 * B_HALF);
                        pg_tma_load_2d_2sm(sB + stage * B_HALF, &tmapB, kc * 64, crow, l_b_full + 8 * stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && lane == 0) {
            uint32_t stage = 0, phase = 0, a_phase = 0;
            uint32_t job = 0;
            for (int64_t u = pair; u < n_units; u += npairs) {
                const int32_t range = (int32_t)(u / p.RB);
                const int64_t c0 = p.goff[p.range_g[range]], c1 = p.goff[p.range_g[range + 1]];
                if (c1 <= c0) continue;
                pg_mbar_wait(bar_a_full, a_phase);
                a_phase ^= 1;
                pg_fence_after();
                const int64_t nchunks = (c1 - c0 + NC - 1) / NC;
                for (int64_t j = 0; j < nchunks; ++j) {
                    const uint32_t stage0 = stage, phase0 = phase;
#pragma unroll
                    for (int rt = 0; rt < MT; ++rt, ++job) {
                        const uint32_t slot = job % NSLOT, it = job / NSLOT;
                        pg_mbar_wait(bar_t_empty + 8 * slot, (it & 1u) ^ 1u);
                        pg_fence_after();
                        const uint32_t td = tmem_base + slot * NC;
                        uint32_t st = stage0, ph = phase0;
#pragma unroll 1
                        for (int kc = 0; kc < KCH; ++kc) {
                            if (rt == 0) {
                                pg_mbar_wait(bar_b_full + 8 * st, ph);
                                pg_fence_after();
                            }
                            const uint64_t db = pg_make_desc(sB + st * B_HALF);
                            const uint64_t da = pg_make_desc(sA + (rt * KCH + kc) * A_TILE);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                pg_mma_bf16_2sm(td, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC, (kc | kk) != 0 ? 1u : 0u);
                            if (rt == MT - 1) pg_commit_2sm(bar_b_empty + 8 * st);
                            if (++st == STAGES) { st = 0; ph ^= 1; }
                        }
                        pg_commit_2sm(bar_t_full + 8 * slot);
                        if (rt == MT - 1) { stage = st; phase = ph; }
                    }
                }
                pg_commit_2sm(bar_a_empty);
            }
        }
    } else {
        // ================= epilogue (both CTAs): warpgroup wg <-> row tile (MT == 2); thread <-> one of this CTA's
        //                    128 bank rows of that tile =================
        const int wq = warp & 3;
        const int wg = (warp - 2) >> 2;
        const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
        uint32_t jobbase = 0;
        for (int64_t u = pair; u < n_units; u += npairs) {
            const int32_t range = (int32_t)(u / p.RB), rb = (int32_t)(u - (int64_t)range * p.RB);
            const int32_t g_lo = p.range_g[range], g_hi = p.range_g[range + 1];
            const int64_t c0 = p.goff[g_lo], c1 = p.goff[g_hi];
            if (c1 <= c0) continue;
            const int rt = MT == 2 ? wg : 0;
            const int64_t tile128 = ((int64_t)rb * MT + rt) * 2 + crank;          // index of this CTA's 128-row tile
            const int64_t row = tile128 * 128 + wq * 32 + lane;
            const int64_t subbase = tile128 * 4 + wq;
            PgPool st;
            pg_pool_init(st, p, g_lo, g_hi, c0);
            const int64_t nchunks = (c1 - c0 + NC - 1) / NC;
            for (int64_t j = 0; j < nchunks; ++j) {
                const uint32_t job = jobbase + (uint32_t)j * MT + rt;
                const uint32_t slot = job % NSLOT, it = job / NSLOT;
                pg_mbar_wait(bar_t_full + 8 * slot, it & 1u);
                pg_fence_after();
                pg_epi_job<NC>(st, p, tmem_base + lane_base + slot * NC, c0 + j * NC, c1, g_hi, (int64_t)p.RB * MT * 8, subbase, lane, row);
                pg_fence_before();
                __syncwarp();
                if (lane == 0) pg_mbar_arrive_cluster(l_t_empty + 8 * slot);
            }
            jobbase += (uint32_t)nchunks * MT;
        }
    }

    // ---- teardown: nobody leaves (or frees TMEM) while the pair still signals each other ----
    pg_fence_before();
    __syncthreads();
    pg_cluster_sync();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- column ranges: range i = groups whose first column falls in [i*T, (i+1)*T) of this batch ----
__global__ void k_pg_ranges(const int64_t* __restrict__ goff, int32_t g_a, int32_t g_b, int64_t T, int32_t n_ranges,
                            int32_t* __restrict__ range_g) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_ranges) return;
    if (i == n_ranges) { range_g[i] = g_b; return; }
    const int64_t target = goff[g_a] + (int64_t)i * T;
    int32_t lo = g_a, hi = g_b;                      // first g in [g_a, g_b] with goff[g] >= target
    while (lo < hi) {
        int32_t mid = lo + (hi - lo) / 2;
        if (goff[mid] >= target) hi = mid; else lo = mid + 1;
    }
    range_g[i] = lo;
}

// ---- merge of the per-(row block, warp) candidate slots of one label group ------------------------
// Picks the `ncand` rows with the largest approximate score and the upper bound on the approximate score of every row
// that is NOT in the list.  With no threshold every slot is full (half the bank per label group), so the scan is
// organised to touch the slots twice: pass A takes, per group of threads, the largest key of its sub-slots; the ncand-th
// largest of those 32 / 64 group maxima is a lower bound T0 on the ncand-th largest key overall.  Pass B compacts the few entries
// >= T0 into shared memory, where they are ranked (descending key, ties by position).  An overfull compaction
// (many equal keys) falls back to a radix select over the slots.
#define PG_MERGE_THREADS 1024
#define PG_MERGE_CAP 2048
__device__ __forceinline__ void pg_load_slot(const float* __restrict__ vals, int c, float (&v)[PG_CS]) {
    const float4* p = reinterpret_cast<const float4*>(vals);
#pragma unroll
    for (int q = 0; q < PG_CS / 4; ++q) {
        float4 t = (q * 4 < c) ? __ldg(p + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
    }
}
__global__ void __launch_bounds__(PG_MERGE_THREADS)
k_pg_merge(const int64_t* __restrict__ goff, int32_t g_base, int32_t nsub, const int32_t* __restrict__ sorted_group,
           const int32_t* __restrict__ glist, const int32_t* __restrict__ group_col,
           const int32_t* __restrict__ slot_cnt, const int32_t* __restrict__ slot_row, const float* __restrict__ slot_val,
           const float* __restrict__ slot_bound, float tau, int32_t ncand, int32_t* __restrict__ cand_row,
           float* __restrict__ gbound, float* __restrict__ cand_val /* may be null: stage-A score of every candidate (diagnostics) */) {
    __shared__ unsigned int s_key[PG_MERGE_CAP];      // pass A: thread maxima (first 1024); pass B: compacted keys
    __shared__ int32_t s_row[PG_MERGE_CAP];
    __shared__ int hist[256];
    __shared__ int s_total, s_digit, s_need, s_out, s_ties, s_m;
    __shared__ unsigned int s_bound_key, s_t0;
    const int tid = threadIdx.x;
    // slot index gl -> label group; with a group list (second-chance merge of the groups whose certificate failed) the
    // launch index selects the group and the compact output row
    // (slots indexed by accumulator column: group_col maps the group to its last column)
    const int gl = glist ? (group_col ? group_col[glist[blockIdx.x]] : glist[blockIdx.x]) - g_base : blockIdx.x;
    const int g = glist ? glist[blockIdx.x] : (sorted_group ? sorted_group[g_base + gl] : g_base + gl);
    if (g < 0) return;
    int32_t* out = cand_row + (int64_t)(glist ? blockIdx.x : g) * ncand;
    float* outv = cand_val ? cand_val + (int64_t)(glist ? blockIdx.x : g) * ncand : nullptr;
    for (int i = tid; i < ncand; i += blockDim.x) out[i] = -1;
    if (goff[g + 1] <= goff[g]) { if (tid == 0) gbound[g] = -3.0e38f; return; }
    const int32_t* cnt = slot_cnt + (int64_t)gl * nsub;
    const int32_t* rows = slot_row + (int64_t)gl * nsub * PG_CS;
    const float* vals = slot_val + (int64_t)gl * nsub * PG_CS;
    const float* bnd = slot_bound + (int64_t)gl * nsub;
    if (tid == 0) { s_total = 0; s_out = 0; s_ties = 0; s_m = 0; s_t0 = 0u; s_bound_key = sdk_fkey(tau); }
    __syncthreads();
    // ---- pass A ----
    int mytotal = 0;
    unsigned int mybound = 0, mymax = 0;
    for (int s = tid; s < nsub; s += blockDim.x) {
        const int c0 = cnt[s];
        const int c = c0 < PG_CS ? c0 : PG_CS;
        mytotal += c;
        if (c0 > PG_CS) { unsigned int kb = sdk_fkey(bnd[s]); mybound = kb > mybound ? kb : mybound; }
        if (c > 0) {
            float v[PG_CS];
            pg_load_slot(vals + (int64_t)s * PG_CS, c, v);
#pragma unroll
            for (int i = 0; i < PG_CS; ++i) { const unsigned int k = i < c ? sdk_fkey(v[i]) : 0u; mymax = k > mymax ? k : mymax; }
        }
    }
    if (mytotal) atomicAdd(&s_total, mytotal);
    if (mybound) atomicMax(&s_bound_key, mybound);
    s_key[tid] = mymax;
    __syncthreads();
    const int total = s_total;
    if (total <= ncand) {
        for (int s = tid; s < nsub; s += blockDim.x) {
            const int c0 = cnt[s];
            const int c = c0 < PG_CS ? c0 : PG_CS;
            for (int i = 0; i < c; ++i) {
                const int pos = atomicAdd(&s_out, 1);
                out[pos] = rows[(int64_t)s * PG_CS + i];
                if (outv) outv[pos] = vals[(int64_t)s * PG_CS + i];
            }
        }
        __syncthreads();
        if (tid == 0) gbound[g] = sdk_funkey(s_bound_key);
        return;
    }
    // T0 = ncand-th largest of NG group maxima (NG = 32 or 64 >= ncand groups of consecutive threads; every group maximum
    // is a distinct entry, so at least ncand entries are >= T0; 0 when fewer than ncand groups saw an entry)
    {
        const int NG = ncand <= 32 ? 32 : 64;
        const int W = (int)blockDim.x / NG;                 // 4..32 threads per group, a power of two inside one warp
        unsigned int gm = mymax;
        for (int off = W >> 1; off >= 1; off >>= 1) {
            const unsigned int o = __shfl_xor_sync(0xffffffffu, gm, off);
            gm = o > gm ? o : gm;
        }
        __syncthreads();                                    // all thread maxima were published above; reuse s_key[0..NG)
        if ((tid & (W - 1)) == 0) s_key[tid / W] = gm;
        __syncthreads();
        if (tid < NG) {
            const unsigned int mine = s_key[tid];
            int rank = 0;
            for (int j = 0; j < NG; ++j) { const unsigned int o = s_key[j]; rank += (o > mine || (o == mine && j < tid)) ? 1 : 0; }
            if (rank == ncand - 1) s_t0 = mine;
        }
    }
    __syncthreads();
    const unsigned int T0 = s_t0;
    __syncthreads();                        // everyone has read T0 / the maxima before s_key is reused
    // ---- pass B: compact the entries >= T0 ----
    for (int s = tid; s < nsub; s += blockDim.x) {
        const int c0 = cnt[s];
        const int c = c0 < PG_CS ? c0 : PG_CS;
        if (c > 0) {
            float v[PG_CS];
            pg_load_slot(vals + (int64_t)s * PG_CS, c, v);
#pragma unroll
            for (int i = 0; i < PG_CS; ++i) {
                if (i < c) {
                    const unsigned int k = sdk_fkey(v[i]);
                    if (k >= T0) {
                        const int pos = atomicAdd(&s_m, 1);
                        if (pos < PG_MERGE_CAP) { s_key[pos] = k; s_row[pos] = rows[(int64_t)s * PG_CS + i]; }
                    }
                }
            }
        }
    }
    __syncthreads();
    const int M = s_m;
    if (M <= PG_MERGE_CAP) {
        // rank inside the compacted list: descending key, ties by bank row (deterministic candidate order)
        for (int i = tid; i < M; i += blockDim.x) {
            const unsigned int k = s_key[i];
            const int32_t r = s_row[i];
            int rank = 0;
            for (int j = 0; j < M; ++j) { const unsigned int o = s_key[j]; rank += (o > k || (o == k && s_row[j] < r)) ? 1 : 0; }
            if (rank < ncand) { out[rank] = r; if (outv) outv[rank] = sdk_funkey(k); }
            if (rank == ncand - 1) s_t0 = k;            // the ncand-th largest key: everything dropped is <= it
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned int T = s_t0;
            gbound[g] = sdk_funkey(s_bound_key > T ? s_bound_key : T);
        }
        return;
    }
    // ---- fallback: radix select of the ncand-th largest key over the slots (4 x 8 bits) ----
    const int64_t flat = (int64_t)nsub * PG_CS;
    unsigned int prefix = 0, mask = 0;
    int need = ncand;
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int64_t f = tid; f < flat; f += blockDim.x) {
            int s = (int)(f / PG_CS), i = (int)(f - (int64_t)s * PG_CS);
            int c = cnt[s];
            if (i < (c < PG_CS ? c : PG_CS)) {
                unsigned int key = sdk_fkey(vals[f]);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, d = 255;
            for (; d > 0; --d) {
                if (cum + hist[d] >= need) break;
                cum += hist[d];
            }
            s_digit = d;
            s_need = need - cum;
        }
        __syncthreads();
        prefix |= ((unsigned int)s_digit) << shift;
        mask |= 255u << shift;
        need = s_need;
        __syncthreads();
    }
    const unsigned int T = prefix;      // ncand-th largest key; take all keys > T and `need` of the keys == T
    for (int64_t f = tid; f < flat; f += blockDim.x) {
        int s = (int)(f / PG_CS), i = (int)(f - (int64_t)s * PG_CS);
        int c = cnt[s];
        if (i < (c < PG_CS ? c : PG_CS)) {
            unsigned int key = sdk_fkey(vals[f]);
            if (key > T) { const int pos = atomicAdd(&s_out, 1); out[pos] = rows[f]; if (outv) outv[pos] = vals[f]; }
            else if (key == T) {
                int t = atomicAdd(&s_ties, 1);
                if (t < need) { const int pos = atomicAdd(&s_out, 1); out[pos] = rows[f]; if (outv) outv[pos] = vals[f]; }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        unsigned int kb = s_bound_key > T ? s_bound_key : T;   // dropped rows are <= T (or below a slot bound / tau)
        gbound[g] = sdk_funkey(kb);
    }
}

// ---- host side -----------------------------------------------------------------------------------
void pg_launch_merge(sdk_ctx* c, const int64_t* d_goff, int32_t g_base, int32_t ngroups, int32_t nsub, const int32_t* d_sorted_group,
                     float tau, int32_t ncand, int32_t* d_cand_row, float* d_gbound, const int32_t* d_glist, const int32_t* d_group_col) {
    if (ngroups <= 0) return;
    sdk_prof_scope ps(c, "merge");
    // small banks (a few hundred sub-slots per label group, tens of thousands of groups): a 256-thread CTA per group
    const int threads = nsub <= 1024 ? 256 : PG_MERGE_THREADS;
    // the first-chance lists also record the stage-A score of every candidate (sdk_stage_a_fetch: measured certificate margin)
    float* cv = (!d_glist && d_cand_row == (int32_t*)c->cand_row.p && c->cand_val.p) ? (float*)c->cand_val.p : nullptr;
    k_pg_merge<<<(unsigned)ngroups, threads, 0, c->stream>>>(d_goff, g_base, nsub, d_sorted_group, d_glist, d_group_col, (const int32_t*)c->slot_cnt.p,
                                                         (const int32_t*)c->slot_row.p, (const float*)c->slot_val.p,
                                                         (const float*)c->slot_bound.p, tau, ncand, d_cand_row, d_gbound, cv);
    c->launches++;
}

struct pg_cfg { int KCH, MT, NC, STAGES; };
// (KCH, MT, NC, STAGES) per padded dimension; shared memory = MT*KCH*16 KB (bank tiles) + STAGES*NC*128 B (ring)
// D <= 192: four 128-column accumulator slots (two per row tile) instead of two 256-column ones.  With one slot per row
// tile the MMAs of chunk j+1 wait for the complete read-out of chunk j, and the read-out runs at the TMEM -> register rate
// (~51-64 B/clk/SM measured: 128 KB per job pair = ~2 000-2 500 cycles against 1 536 MMA cycles at D = 192), so the
// pipe idled a third of the time; with two slots per row tile the kernel runs at the read-out rate (config 3 with max
// pooling: 67.1 -> 64.0 ms; generic mean: 68.0 -> 64.4 ms).  SDK_PG_NC=256 restores the wide slots (A/B runs).
static bool pg_nc128() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SDK_PG_NC"); v = (e && atoi(e) == 256) ? 0 : 1; }
    return v == 1;
}
static pg_cfg pg_config_for(int32_t Dp) {
    int kch = Dp / 64;
    if (pg_nc128() && kch <= 3) return {kch, 2, 128, 6};
    switch (kch) {
        case 1: return {1, 2, 256, 6};
        case 2: return {2, 2, 256, 5};
        case 3: return {3, 2, 256, 4};
        case 4: return {4, 2, 128, 6};
        case 5: return {5, 1, 256, 4};
        case 6: return {6, 1, 256, 4};
        case 7: return {7, 1, 256, 3};
        default: return {8, 1, 256, 3};
    }
}

int sdk_poolgemm_supported(int32_t Dp) { return Dp >= 64 && Dp <= 512 && Dp % 64 == 0; }

template <int KCH, int MT, int NC, int STAGES>
static int pg_launch_t(sdk_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const PgParams& p, int grid) {
    constexpr size_t smem = (size_t)MT * KCH * 16384 + (size_t)STAGES * NC * 128 + 256 + 1024;
    static_assert(smem <= PG_SMEM_LIMIT, "shared memory budget");
    auto kern = k_poolgemm<KCH, MT, NC, STAGES>;
    SDK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, MT == 2 ? PG_THREADS2 : PG_THREADS, smem, c->stream>>>(ta, tb, p);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

static int pg_launch(sdk_ctx* c, const pg_cfg& cfg, const CUtensorMap& ta, const CUtensorMap& tb, const PgParams& p, int grid) {
    if (cfg.NC == 128 && cfg.KCH <= 3) {
        switch (cfg.KCH) {
            case 1: return pg_launch_t<1, 2, 128, 6>(c, ta, tb, p, grid);
            case 2: return pg_launch_t<2, 2, 128, 6>(c, ta, tb, p, grid);
            default: return pg_launch_t<3, 2, 128, 6>(c, ta, tb, p, grid);
        }
    }
    switch (cfg.KCH) {
        case 1: return pg_launch_t<1, 2, 256, 6>(c, ta, tb, p, grid);
        case 2: return pg_launch_t<2, 2, 256, 5>(c, ta, tb, p, grid);
        case 3: return pg_launch_t<3, 2, 256, 4>(c, ta, tb, p, grid);
        case 4: return pg_launch_t<4, 2, 128, 6>(c, ta, tb, p, grid);
        case 5: return pg_launch_t<5, 1, 256, 4>(c, ta, tb, p, grid);
        case 6: return pg_launch_t<6, 1, 256, 4>(c, ta, tb, p, grid);
        case 7: return pg_launch_t<7, 1, 256, 3>(c, ta, tb, p, grid);
        default: return pg_launch_t<8, 1, 256, 3>(c, ta, tb, p, grid);
    }
}

// (KCH, MT, NC, STAGES) of the 2-CTA kernel; per CTA: MT*KCH*16 KB of bank tiles + STAGES*(NC/2)*128 B ring
static pg_cfg pg_config2_for(int32_t Dp) {
    int kch = Dp / 64;
    switch (kch) {
        case 1: return {1, 2, 256, 8};
        case 2: return {2, 2, 256, 8};
        case 3: return {3, 2, 256, 8};
        case 4: return {4, 2, 256, 6};
        case 5: return {5, 1, 256, 8};
        case 6: return {6, 1, 256, 8};
        case 7: return {7, 1, 256, 6};
        default: return {8, 1, 256, 6};
    }
}

template <int KCH, int MT, int NC, int STAGES>
static int pg_launch2_t(sdk_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const PgParams& p, int grid) {
    constexpr size_t smem = (size_t)MT * KCH * 16384 + (size_t)STAGES * (NC / 2) * 128 + 256 + 1024;
    static_assert(smem <= PG_SMEM_LIMIT, "shared memory budget");
    auto kern = k_poolgemm2<KCH, MT, NC, STAGES>;
    SDK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, MT == 2 ? PG_THREADS2 : PG_THREADS, smem, c->stream>>>(ta, tb, p);      // __cluster_dims__(2,1,1): grid must be even
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

static int pg_launch2(sdk_ctx* c, const pg_cfg& cfg, const CUtensorMap& ta, const CUtensorMap& tb, const PgParams& p, int grid) {
    switch (cfg.KCH) {
        case 1: return pg_launch2_t<1, 2, 256, 8>(c, ta, tb, p, grid);
        case 2: return pg_launch2_t<2, 2, 256, 8>(c, ta, tb, p, grid);
        case 3: return pg_launch2_t<3, 2, 256, 8>(c, ta, tb, p, grid);
        case 4: return pg_launch2_t<4, 2, 256, 6>(c, ta, tb, p, grid);
        case 5: return pg_launch2_t<5, 1, 256, 8>(c, ta, tb, p, grid);
        case 6: return pg_launch2_t<6, 1, 256, 8>(c, ta, tb, p, grid);
        case 7: return pg_launch2_t<7, 1, 256, 6>(c, ta, tb, p, grid);
        default: return pg_launch2_t<8, 1, 256, 6>(c, ta, tb, p, grid);
    }
}

// Columns per range.  Units (range x row block) are dealt to the persistent CTAs (or CTA pairs) round robin, so the step
// takes ceil(units / workers) unit times: the number of ranges is chosen to waste little of the last wave while keeping
// the per-unit bank-tile reload (worth ~384 columns of MMA time) amortised.  Whole groups, multiples of NC.
static int64_t pg_range_cols(int64_t ncols, int32_t RB, int NC, int64_t workers) {
    const double reload = 384.0;
    int64_t best_n = 1;
    double best = 1e300;
    const int64_t n_max = std::max<int64_t>(1, std::min<int64_t>(64, ncols / (4 * NC)));
    for (int64_t n = 1; n <= n_max; ++n) {
        const int64_t waves = (n * RB + workers - 1) / workers;
        const double cost = (double)waves * ((double)ncols / (double)n + reload);
        if (cost < best * 0.999) { best = cost; best_n = n; }
    }
    int64_t T = (ncols + best_n - 1) / best_n;
    T = (T + NC - 1) / NC * NC;
    if (T < 4 * NC) T = 4 * NC;
    return T;
}

static int pg_run(sdk_ctx* c, const __nv_bfloat16* d_rows, int64_t P, const __nv_bfloat16* d_cols, int64_t N, int32_t Dp,
                  const int64_t* d_goff, int32_t G, int32_t pool, int32_t mode, float tau, int32_t ncand, int32_t* d_cand_row,
                  float* d_gbound, float* d_dense) {
    if (!sdk_poolgemm_supported(Dp) || !c->tmap_encode) return sdk_fail(c, SDK_EINVAL, "tcgen05 path unavailable");
    if (N > 0x7fffffffLL || P > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "tcgen05 path: at most 2^31-1 rows / segments");
    const bool two = c->opt_cta_group == 2 && c->sm_count >= 2;
    const pg_cfg cfg = two ? pg_config2_for(Dp) : pg_config_for(Dp);
    const int64_t rows_per_block = (int64_t)cfg.MT * (two ? 256 : 128);
    const int32_t RB = (int32_t)((P + rows_per_block - 1) / rows_per_block);
    CUtensorMap ta, tb;
    SDK_TRY(pg_make_tmap(c, &ta, d_rows, P, Dp, 128));
    SDK_TRY(pg_make_tmap(c, &tb, d_cols, N, Dp, (uint32_t)(two ? cfg.NC / 2 : cfg.NC)));
    // group offsets are needed on the host only to batch groups; read them once (G+1 int64)
    std::vector<int64_t> hgoff((size_t)G + 1);
    SDK_CUDA(c, cudaMemcpyAsync(hgoff.data(), d_goff, ((size_t)G + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    // batches of label groups so that the candidate slots stay under ~6 GB
    const int32_t nsub = (int32_t)(RB * rows_per_block / 32);   // one sub-slot per 32 bank rows (warp) per label group
    const size_t per_group = (size_t)nsub * (PG_CS * 8 + 8);
    int64_t gbatch = mode == 0 ? (int64_t)((size_t)(6144ull << 20) / per_group) : (int64_t)G;
    if (gbatch < 1) gbatch = 1;
    if (gbatch > G) gbatch = G;
    if (mode == 0) {
        SDK_TRY(sdk_reserve(c, c->slot_cnt, (size_t)gbatch * nsub * 4));
        SDK_TRY(sdk_reserve(c, c->slot_bound, (size_t)gbatch * nsub * 4));
        SDK_TRY(sdk_reserve(c, c->slot_row, (size_t)gbatch * nsub * PG_CS * 4));
        SDK_TRY(sdk_reserve(c, c->slot_val, (size_t)gbatch * nsub * PG_CS * 4));
    }
    const int64_t workers = two ? c->sm_count / 2 : c->sm_count;
    c->slot_g0 = c->slot_g1 = 0;
    for (int64_t ga = 0; ga < G; ga += gbatch) {
        const int64_t gb = std::min<int64_t>(G, ga + gbatch);
        const int64_t ncols = hgoff[gb] - hgoff[ga];
        if (ncols > 0) {
            const int64_t T = pg_range_cols(ncols, RB, cfg.NC, workers);
            const int32_t n_ranges = (int32_t)((ncols + T - 1) / T);
            SDK_TRY(sdk_reserve(c, c->range_g, (size_t)(n_ranges + 1) * 4));
            k_pg_ranges<<<(n_ranges + 1 + 255) / 256, 256, 0, c->stream>>>(d_goff, (int32_t)ga, (int32_t)gb, T, n_ranges,
                                                                          (int32_t*)c->range_g.p);
            c->launches++;
            SDK_CUDA(c, cudaGetLastError());
            PgParams p;
            p.goff = d_goff;
            p.range_g = (const int32_t*)c->range_g.p;
            p.n_ranges = n_ranges;
            p.RB = RB;
            p.P = P;
            p.g_base = (int32_t)ga;
            p.pool = pool;
            p.tau = tau;
            p.mode = mode;
            p.slot_cnt = (int32_t*)c->slot_cnt.p;
            p.slot_row = (int32_t*)c->slot_row.p;
            p.slot_val = (float*)c->slot_val.p;
            p.slot_bound = (float*)c->slot_bound.p;
            p.dense_out = d_dense;
            p.dense_ld = G;
            p.nsub = nsub;
            p.kth = nullptr;
            if (mode == 0 && c->kth_on) {
                SDK_TRY(sdk_reserve(c, c->kth, (size_t)gbatch * PG_KTH * 4));
                SDK_CUDA(c, cudaMemsetAsync(c->kth.p, 0, (size_t)(gb - ga) * PG_KTH * 4, c->stream));
                p.kth = (uint32_t*)c->kth.p;
            }
            const int64_t n_units = (int64_t)n_ranges * RB;
            {
                sdk_prof_scope ps(c, "poolgemm");
                if (two) {
                    const int grid = 2 * (int)std::min<int64_t>(n_units, workers);
                    SDK_TRY(pg_launch2(c, cfg, ta, tb, p, grid));
                } else {
                    const int grid = (int)std::min<int64_t>(n_units, workers);
                    SDK_TRY(pg_launch(c, cfg, ta, tb, p, grid));
                }
            }
        }
        if (mode == 0) {
            pg_launch_merge(c, d_goff, (int32_t)ga, (int32_t)(gb - ga), nsub, nullptr, tau, ncand, d_cand_row, d_gbound);
            SDK_CUDA(c, cudaGetLastError());
            c->slot_by_col = false;
            c->slot_g0 = (int32_t)ga;        // groups whose candidate slots are still in memory after the call
            c->slot_g1 = (int32_t)gb;
            c->slot_nsub = nsub;
        }
    }
    return SDK_OK;
}

int sdk_launch_poolgemm_candidates(sdk_ctx* c, const __nv_bfloat16* d_bank, int64_t P, const __nv_bfloat16* d_seg, int64_t N,
                                   int32_t Dp, const int64_t* d_goff, int32_t G, int32_t pool, float tau, int32_t ncand,
                                   int32_t* d_cand_row, float* d_gbound) {
    return pg_run(c, d_bank, P, d_seg, N, Dp, d_goff, G, pool, 0, tau, ncand, d_cand_row, d_gbound, nullptr);
}

int sdk_launch_poolgemm_dense(sdk_ctx* c, const __nv_bfloat16* d_rows, int64_t P, const __nv_bfloat16* d_cols, int64_t N,
                              int32_t Dp, const int64_t* d_goff, int32_t G, int32_t pool, float* d_out) {
    return pg_run(c, d_rows, P, d_cols, N, Dp, d_goff, G, pool, 1, 0.f, 0, nullptr, nullptr, d_out);
}

// Second chance for label groups whose top-k certificate failed: a wider candidate list from the slots of the last batch.
int sdk_launch_poolgemm_remerge(sdk_ctx* c, const int64_t* d_goff, const int32_t* d_glist, int32_t ngroups, float tau, int32_t ncand,
                                int32_t* d_cand_row /*[ngroups,ncand]*/, float* d_gbound /*[G]*/) {
    if (ngroups <= 0) return SDK_OK;
    pg_launch_merge(c, d_goff, c->slot_g0, ngroups, c->slot_nsub, nullptr, tau, ncand, d_cand_row, d_gbound, d_glist,
                    c->slot_by_col ? (const int32_t*)c->pa_col_last.p : nullptr);
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}
