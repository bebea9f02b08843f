// gemv.cu -- K2d: stage A for a HANDFUL of query segments against a big bank ("8 pooled label centroids vs a 125k-row
// shard", BASELINE config 4 variant (i)).  One pass over the bank, HBM-bound: 2*Dp bytes per bank row, nothing else.
//
// The tcgen05 kernels need at least a 256-column accumulator tile and reload a 128-row bank tile per unit without
// overlap; for <= 8 query segments that is 32x wasted tensor work and an exposed TMA round trip per 128 rows.  Here the
// bank is streamed through shared memory by TMA (one elected producer thread, 2-D boxes of 64 rows x 64 bf16 with the
// 128-byte swizzle): each of the eight consumer warps owns a private ring of 3 stages (mbarrier full/empty pairs; a
// barrier is only ever waited on by its own warp, phase after phase), 192 KB in flight per SM (sixteen consumer warps,
// 32-row stages).  The arithmetic is the
// one tensor-core shape that fits 8 queries exactly: mma.sync m16n8k16 (bf16 in, fp32 accumulate) -- bank rows are the
// M side (ldmatrix.x4 straight from the swizzled tile, conflict-free), the 8 queries are the N side (fragments pre-packed
// in shared memory, 2 words per lane and 16-wide K step), 8 accumulator registers hold a 32-row chunk across the K
// chunks.  That is ~50 instructions per 32x64 stage, so the SM only waits for HBM.  After the last K chunk the
// accumulator fragments are transposed through shared memory; lane i then owns rows i and i+32, pools their scores per
// label group and flushes the candidate slots exactly like the tensor-core epilogues (tcgen05.cuh: pg_flush), so merge /
// canonical re-score / certificate downstream are unchanged.  (tcgen05.mma needs N >= 16 and a TMEM round trip; for
// an HBM-bound 8-column product the warp-level MMA is the better fit.)
// Products of bf16 operands are exact in fp32; only the fp32 accumulation order differs from the canonical arithmetic,
// well inside the certificate's eps.
#include "tcgen05.cuh"

#define GV_NQ 8
#define GV_CW 16                     // consumer warps
#define GV_THREADS (32 * (GV_CW + 1))
#define GV_RING 3                    // stages per consumer warp
#define GV_STAGES (GV_CW * GV_RING)
#define GV_ROWS 32                   // bank rows per chunk / stage (one candidate sub-slot)
#define GV_STAGE_BYTES 4096u         // 32 rows x 64 bf16

struct GvParams {
    PgParams pg;
    int32_t N, G, Dp;
    int32_t nsub;                 // sub-slots per label group = ceil(P / 32)
    int32_t nchunk;               // 32-row chunks of the bank (== nsub)
};

// Pools and flushes one 32-row chunk (one sub-slot per label group).  Rolled loops: every warp runs this only a few
// times, so instruction-cache footprint matters more than unrolling; inlined so that the shared-memory operands are
// read with shared-space loads (a generic-pointer version stalled on every label).
static __device__ __forceinline__ void gv_flush_chunk(const GvParams& q, const float* __restrict__ sc /*[32][GV_NQ]*/, const int32_t* s_grp,
                                                   const int32_t* s_len, int64_t chunk, int lane) {
    const PgParams& p = q.pg;
    const int64_t row = chunk * GV_ROWS + lane;
    const float* v = sc + lane * GV_NQ;
    // running k-th best thresholds of the (at most GV_NQ) label groups: all loads are issued before the first is used --
    // one L2 round trip per chunk instead of one per label (the flushes below then skip the per-flush lookup)
    // (lane s keeps the threshold of query slot s; the flush loop stays rolled -- every warp runs it only a few times, so
    //  instruction-cache footprint matters more than unrolling)
    uint32_t thr_mine = 0u;
    if (p.kth) {
        uint32_t ka[GV_NQ], kb[GV_NQ];
#pragma unroll
        for (int s = 0; s < GV_NQ; ++s) {
            const int g = s < q.N ? s_grp[s] : -1;
            const uint32_t* kp = p.kth + (int64_t)(g < 0 ? 0 : g - p.g_base) * PG_KTH;
            ka[s] = g >= 0 ? __ldcg(kp + lane) : 0u;
            kb[s] = g >= 0 ? __ldcg(kp + lane + 32) : 0u;
        }
#pragma unroll
        for (int s = 0; s < GV_NQ; ++s) {
            const uint32_t t = __reduce_min_sync(0xffffffffu, ka[s] < kb[s] ? ka[s] : kb[s]);
            thr_mine = lane == s ? t : thr_mine;
        }
    }
    float a = p.pool == 0 ? 0.f : -3.0e38f;
#pragma unroll 1
    for (int s = 0; s < q.N; ++s) {
        a = p.pool == 0 ? a + v[s] : fmaxf(a, v[s]);
        const int g = s_grp[s];
        if (s + 1 == q.N || s_grp[s + 1] != g) {
            const float val = p.pool == 0 ? a * (1.0f / (float)s_len[s]) : a;
            const int64_t sg = g - p.g_base, sub = sg * p.nsub + chunk;
            bool pass = (val >= p.tau) && (row < p.P);
            if (p.kth) {
                const uint32_t thr = __shfl_sync(0xffffffffu, thr_mine, s);
                const uint32_t key = sdk_fkey(val);
                pg_kth_update(p, sg, chunk, pass ? key : 0u, thr, lane);
                pass = pass && key >= thr;
            }
            const uint32_t mpass = __ballot_sync(0xffffffffu, pass);
            if (lane == 0 && mpass == 0) p.slot_cnt[sub] = 0;
            if (mpass != 0) pg_flush_write(p, val, pass, mpass, sub, lane, row);
            a = p.pool == 0 ? 0.f : -3.0e38f;
        }
    }
}

template <int KCH>
__global__ void __launch_bounds__(GV_THREADS, 1)
k_gemv8(const __grid_constant__ CUtensorMap tmapBank, const uint32_t* __restrict__ seg /*[N, Dp/2] bf16x2*/, const __grid_constant__ GvParams q) {
    extern __shared__ uint8_t gv_smem_raw[];
    __shared__ __align__(16) float s_c[GV_CW][GV_ROWS][GV_NQ]; // accumulator transpose: [row of the chunk][query]
    __shared__ __align__(8) uint2 s_bq[KCH * 4][32];           // B fragments of every 16-wide K step, per lane
    __shared__ int32_t s_grp[GV_NQ];              // label group of query s (-1: padding)
    __shared__ int32_t s_len[GV_NQ];              // segments of that group
    __shared__ __align__(8) uint64_t s_bar[2 * GV_STAGES];
    const PgParams& p = q.pg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = q.Dp >> 1;
    const uint32_t raw = pg_smem_u32(gv_smem_raw);
    const uint32_t ring = (raw + 1023u) & ~1023u;                   // the swizzle atom is 1024 bytes
    const uint32_t bar_full = pg_smem_u32(s_bar), bar_empty = bar_full + 8 * GV_STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GV_STAGES; ++s) { pg_mbar_init(bar_full + 8 * s, 1); pg_mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + GV_NQ) {
        const int s = threadIdx.x - 32;
        int g = -1, len = 0;
        if (s < q.N) {
            int lo = 0, hi = q.G;                 // last g with goff[g] <= s (the non-empty group that holds segment s)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (p.goff[mid] <= s) lo = mid; else hi = mid;
            }
            g = lo;
            len = (int)(p.goff[g + 1] - p.goff[g]);
        }
        s_grp[s] = g;
        s_len[s] = len;
    }
    // B fragments (mma.sync m16n8k16 "col" operand): lane l holds query n = l / 4, k = 16 * ks + 2 * (l % 4) (+8).  Kept in
    // shared memory so that the K loop stays rolled: every warp runs the loop body only a couple of times, and a fully
    // unrolled kernel spent more time missing the instruction cache than computing
    for (int i = threadIdx.x; i < KCH * 4 * 32; i += GV_THREADS) {
        const int ks = i >> 5, l = i & 31, n = l >> 2, kw = l & 3;
        uint2 b = make_uint2(0u, 0u);
        if (n < q.N) {
            b.x = __ldg(seg + (int64_t)n * pitch + ks * 8 + kw);
            b.y = __ldg(seg + (int64_t)n * pitch + ks * 8 + 4 + kw);
        }
        s_bq[ks][l] = b;
    }
    __syncthreads();
    // chunks of this CTA: blockIdx.x, + gridDim.x, ...; the k-th goes to consumer warp k % GV_CW as that warp's chunk
    // number n = k / GV_CW; its K chunk kc is the warp's stage number m = KCH * n + kc -> stage (m % GV_RING) of its ring
    const int64_t n_mine = q.nchunk > (int64_t)blockIdx.x ? (q.nchunk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (warp == 0) {
        if (lane == 0) {
            // issue order: K chunk by K chunk, round robin over the warps, so that a full ring of one warp never holds
            // back the loads of the others
            for (int64_t kb = 0; kb < n_mine; kb += GV_CW) {
                const int64_t n = kb / GV_CW;
                for (int kc = 0; kc < KCH; ++kc) {
                    for (int w = 0; w < GV_CW; ++w) {
                        const int64_t k = kb + w;
                        if (k >= n_mine) break;
                        const int64_t chunk = blockIdx.x + k * gridDim.x;
                        const int64_t m = n * KCH + kc;
                        const int st = w * GV_RING + (int)(m % GV_RING);
                        const uint32_t use = (uint32_t)(m / GV_RING);
                        pg_mbar_wait(bar_empty + 8 * st, (use & 1u) ^ 1u);
                        pg_mbar_expect_tx(bar_full + 8 * st, GV_STAGE_BYTES);
                        pg_tma_load_2d(ring + st * GV_STAGE_BYTES, &tmapBank, kc * 64, (int32_t)(chunk * GV_ROWS), bar_full + 8 * st);   // rows past P: zero fill
                    }
                }
            }
        }
        return;
    }
    const int cw = warp - 1;
    // ldmatrix.x4 source row of this lane inside a 16-row tile, and which 8-wide K half it addresses
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhalf = lane >> 4;
    int64_t m = 0;                                  // this warp's stage counter (walks its private ring)
    for (int64_t k = cw; k < n_mine; k += GV_CW) {
        const int64_t chunk = blockIdx.x + k * gridDim.x;
        float acc[GV_ROWS / 16][4];
#pragma unroll
        for (int rt = 0; rt < GV_ROWS / 16; ++rt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[rt][i] = 0.f;
#pragma unroll 1
        for (int kc = 0; kc < KCH; ++kc, ++m) {
            const int st = cw * GV_RING + (int)(m % GV_RING);
            const uint32_t use = (uint32_t)(m / GV_RING);
            pg_mbar_wait(bar_full + 8 * st, use & 1u);
            const uint32_t tile = ring + st * GV_STAGE_BYTES;
#pragma unroll
            for (int rt = 0; rt < GV_ROWS / 16; ++rt) {
                const int r = rt * 16 + lrow;                       // row of the tile; swizzle: 16-byte piece ^= row % 8
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t addr = tile + r * 128 + (((2 * ks + lhalf) ^ (r & 7)) << 4);
                    uint32_t a0, a1, a2, a3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
                    const uint2 b = s_bq[kc * 4 + ks][lane];
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[rt][0]), "+f"(acc[rt][1]), "+f"(acc[rt][2]), "+f"(acc[rt][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b.x), "r"(b.y));
                }
            }
            __syncwarp();
            if (lane == 0) pg_mbar_arrive(bar_empty + 8 * st);    // the stage may be refilled while this warp goes on
        }
        // transpose: fragment (row = lane/4 (+8), queries 2*(lane%4), +1)  ->  lane <-> row
        __syncwarp();
#pragma unroll
        for (int rt = 0; rt < GV_ROWS / 16; ++rt) {
            const int r = rt * 16 + (lane >> 2), cq = 2 * (lane & 3);
            *reinterpret_cast<float2*>(&s_c[cw][r][cq]) = make_float2(acc[rt][0], acc[rt][1]);
            *reinterpret_cast<float2*>(&s_c[cw][r + 8][cq]) = make_float2(acc[rt][2], acc[rt][3]);
        }
        __syncwarp();
        // lane <-> bank row chunk*32 + lane: pool the queries of each label group (ascending segment order), flush
        gv_flush_chunk(q, &s_c[cw][0][0], s_grp, s_len, chunk, lane);
        __syncwarp();                                               // s_c is rewritten by the next chunk
    }
}

int sdk_gemv_applicable(int64_t N, int32_t Dp) { return N >= 1 && N <= GV_NQ && sdk_poolgemm_supported(Dp); }

// Candidates + bound per label group, like sdk_launch_poolgemm_candidates, for N <= 8 query segments.
int sdk_launch_gemv_candidates(sdk_ctx* c, const __nv_bfloat16* d_bank, int64_t P, const __nv_bfloat16* d_seg, int64_t N, int32_t Dp,
                               const int64_t* d_goff, int32_t G, int32_t pool, float tau, int32_t ncand, int32_t* d_cand_row,
                               float* d_gbound) {
    if (!sdk_gemv_applicable(N, Dp)) return sdk_fail(c, SDK_EINVAL, "gemv path: needs 1..8 segments and a supported D");
    if (P > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "gemv path: at most 2^31-1 bank rows");
    const int32_t nsub = (int32_t)((P + 31) / 32);
    // slots only for groups that can hold a segment: groups are flushed by id, so the arrays span all G groups
    const size_t per_group = (size_t)nsub * (PG_CS * 8 + 8);
    if (per_group * (size_t)G > (8192ull << 20)) return sdk_fail(c, SDK_EINVAL, "gemv path: too many label groups for this bank");
    SDK_TRY(sdk_reserve(c, c->slot_cnt, (size_t)G * nsub * 4));
    SDK_TRY(sdk_reserve(c, c->slot_bound, (size_t)G * nsub * 4));
    SDK_TRY(sdk_reserve(c, c->slot_row, (size_t)G * nsub * PG_CS * 4));
    SDK_TRY(sdk_reserve(c, c->slot_val, (size_t)G * nsub * PG_CS * 4));
    GvParams q;
    q.pg.goff = d_goff;
    q.pg.range_g = nullptr;
    q.pg.n_ranges = 0;
    q.pg.RB = 0;
    q.pg.P = P;
    q.pg.g_base = 0;
    q.pg.pool = pool;
    q.pg.tau = tau;
    q.pg.mode = 0;
    q.pg.slot_cnt = (int32_t*)c->slot_cnt.p;
    q.pg.slot_row = (int32_t*)c->slot_row.p;
    q.pg.slot_val = (float*)c->slot_val.p;
    q.pg.slot_bound = (float*)c->slot_bound.p;
    q.pg.dense_out = nullptr;
    q.pg.dense_ld = 0;
    q.pg.nsub = nsub;
    q.pg.kth = nullptr;
    // (with <= 8 queries the slot volume is small and the per-chunk bucket lookups cost more than they save -- measured
    //  46.1 vs 43.6 us on config 4-i -- so the pruning is only taken when forced: option kth = 2)
    if (c->kth_on && c->opt_kth == 2) {
        SDK_TRY(sdk_reserve(c, c->kth, (size_t)G * PG_KTH * 4));
        SDK_CUDA(c, cudaMemsetAsync(c->kth.p, 0, (size_t)G * PG_KTH * 4, c->stream));
        q.pg.kth = (uint32_t*)c->kth.p;
    }
    q.N = (int32_t)N;
    q.G = G;
    q.Dp = Dp;
    q.nsub = nsub;
    q.nchunk = nsub;
    const int grid = (int)std::min<int64_t>((q.nchunk + GV_CW - 1) / GV_CW, (int64_t)c->sm_count);
    const size_t smem = (size_t)GV_STAGES * GV_STAGE_BYTES + 1024;
    if (!c->tmap_encode) return sdk_fail(c, SDK_EINVAL, "gemv path: cuTensorMapEncodeTiled unavailable");
    CUtensorMap tb;
    SDK_TRY(pg_make_tmap(c, &tb, d_bank, P, Dp, GV_ROWS));
    {
        sdk_prof_scope ps(c, "poolgemm");         // stage A of the certified top-k, whichever kernel runs it
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(d_seg);
#define GV_LAUNCH(K)                                                                                            \
    do {                                                                                                            \
        SDK_CUDA(c, cudaFuncSetAttribute(k_gemv8<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        k_gemv8<K><<<grid, GV_THREADS, smem, c->stream>>>(tb, s32, q);                                             \
    } while (0)
        switch (Dp / 64) {
            case 1: GV_LAUNCH(1); break;
            case 2: GV_LAUNCH(2); break;
            case 3: GV_LAUNCH(3); break;
            case 4: GV_LAUNCH(4); break;
            case 5: GV_LAUNCH(5); break;
            case 6: GV_LAUNCH(6); break;
            case 7: GV_LAUNCH(7); break;
            default: GV_LAUNCH(8); break;
        }
#undef GV_LAUNCH
        c->launches++;
        SDK_CUDA(c, cudaGetLastError());
    }
    pg_launch_merge(c, d_goff, 0, G, nsub, nullptr, tau, ncand, d_cand_row, d_gbound);
    SDK_CUDA(c, cudaGetLastError());
    c->slot_g0 = 0;
    c->slot_g1 = G;
    c->slot_nsub = nsub;
    c->slot_by_col = false;
    return SDK_OK;
}
