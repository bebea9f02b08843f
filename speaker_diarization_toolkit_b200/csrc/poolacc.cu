// poolacc.cu -- K2c: mean pooling fused INTO the tensor-core accumulation ("accumulate-pooling").
//
// poolgemm.cu puts segments on the accumulator columns and pools them in the epilogue: every scored pair is read
// out of TMEM once and added on the CUDA cores.  With D = 192 a 128x256 tile is only 12 MMAs (1536 tensor cycles)
// but 128 KB of TMEM read-out + 32k FADDs, so the epilogue, not the tensor pipe, is the bound (profiles/r01_*).
//
// For MEAN pooling the sum over a label's segments can be done by the MMA itself.  Segments are laid out
// "group interleaved": label groups are sorted by size and cut into blocks of 256 groups; step t of a block is the
// 256-row slab holding the t-th segment of each of its 256 groups (zero rows once a group is exhausted).  Column j of
// the accumulator tile therefore always belongs to group j of the block, and issuing the MMAs of step 0,1,..,T-1
// into the SAME TMEM tile (accumulate = true across steps as well as across K) leaves
//        D[row, j] = sum_t  bank[row] . seg_t(group j)            -- the pooled sum, every pair contracted --
// after T*D/16 MMAs.  The tile is read out once per 256 GROUPS instead of once per 256 segments (~250x fewer TMEM
// reads and epilogue instructions for hour-long recordings), the B ring needs no chunk residency, and the kernel
// behaves like a large-K GEMM.  Sorting by size keeps the zero padding to a few percent.
//
// Everything downstream is unchanged: the epilogue flushes per (group, 32 bank rows) candidate slots, k_pg_merge picks
// the candidates, exact.cu re-scores them canonically (reading segments from the interleaved layout), select.cu
// certifies.  Max pooling, few-group shapes and the dense (config 5) mode stay on poolgemm.cu.
#include "tcgen05.cuh"

#define PA_NB 256                 // label groups per block == accumulator columns
#define PA_BUCKETS 65536

// ---- planning: counting sort of the groups by (clipped) size, descending -------------------------------------
__global__ void k_pa_hist(const int64_t* __restrict__ goff, int32_t G, int32_t* __restrict__ hist) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int64_t n = goff[g + 1] - goff[g];
    int b = PA_BUCKETS - 1 - (int)(n < PA_BUCKETS - 1 ? n : PA_BUCKETS - 1);
    atomicAdd(&hist[b], 1);
}
// exclusive scan of 65536 buckets by one CTA of 1024 threads (64 buckets each)
__global__ void __launch_bounds__(1024) k_pa_scan(int32_t* __restrict__ hist) {
    __shared__ int part[1024];
    const int t = threadIdx.x;
    int loc[64], sum = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) { loc[i] = hist[t * 64 + i]; sum += loc[i]; }
    part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int base = part[t] - sum;
#pragma unroll
    for (int i = 0; i < 64; ++i) { hist[t * 64 + i] = base; base += loc[i]; }
}
__global__ void k_pa_scatter(const int64_t* __restrict__ goff, int32_t G, int32_t* __restrict__ cursor, int32_t* __restrict__ sorted_group,
                             int32_t* __restrict__ group_pos) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int64_t n = goff[g + 1] - goff[g];
    int b = PA_BUCKETS - 1 - (int)(n < PA_BUCKETS - 1 ? n : PA_BUCKETS - 1);
    int pos = atomicAdd(&cursor[b], 1);
    sorted_group[pos] = g;
    group_pos[g] = pos;
}
// per block of 256 sorted groups: T_b = largest group; step0 = exclusive prefix of T_b; plan_total[0] = total steps
__global__ void k_pa_blocks(const int64_t* __restrict__ goff, const int32_t* __restrict__ sorted_group, int32_t G, int32_t n_blocks,
                            int32_t Gpad, int32_t* __restrict__ sorted_pad, int32_t* __restrict__ blockT, int64_t* __restrict__ step0,
                            int64_t* __restrict__ plan_total) {
    // phase 1: T_b
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += gridDim.x * blockDim.x) {
        int64_t m = 0;
        for (int j = 0; j < PA_NB; ++j) {
            int pos = b * PA_NB + j;
            int g = pos < G ? sorted_group[pos] : -1;
            sorted_pad[pos] = g;
            if (g >= 0) { int64_t n = goff[g + 1] - goff[g]; m = n > m ? n : m; }
        }
        blockT[b] = (int32_t)(m > 0x7fffffff ? 0x7fffffff : m);
    }
    (void)Gpad;
    (void)step0;
    (void)plan_total;
}
__global__ void k_pa_steps(const int32_t* __restrict__ blockT, int32_t n_blocks, int64_t* __restrict__ step0, int64_t* __restrict__ plan_total) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int64_t acc = 0;
    for (int b = 0; b < n_blocks; ++b) { step0[b] = acc; acc += blockT[b]; }
    step0[n_blocks] = acc;
    plan_total[0] = acc;
}
// step -> block
__global__ void k_pa_step_table(const int64_t* __restrict__ step0, int32_t n_blocks, int64_t S, int32_t* __restrict__ step_block) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int lo = 0, hi = n_blocks;                 // last b with step0[b] <= s
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (step0[mid] <= s) lo = mid; else hi = mid;
    }
    step_block[s] = lo;
}
// per group: first row and row stride of its segments in the interleaved matrix (for the canonical re-score)
__global__ void k_pa_group_rows(const int32_t* __restrict__ group_pos, const int64_t* __restrict__ step0, int32_t G,
                                int64_t* __restrict__ seg_base) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int pos = group_pos[g];
    seg_base[g] = step0[pos / PA_NB] * PA_NB + (pos % PA_NB);
}

// ---- K1 (gather form): the interleaved bf16 matrix is written in runs of 8 consecutive destination rows per warp ---
// (8 slots of one step: block b and step t are looked up once, lanes 0..7 resolve the 8 source segments in parallel,
// then the 8 rows' 128-bit loads are all in flight before the first norm is reduced -- the per-row metadata chain
// step -> block -> group -> goff would otherwise serialise ~4 dependent loads in front of every row).
// Arithmetic = k_normalize_vec (canonical); padding slots become zero rows.
template <int NQ>
__global__ void __launch_bounds__(256)
k_pa_normalize_gather(const float* __restrict__ x, int32_t D, int32_t Dp, const int64_t* __restrict__ goff,
                      const int32_t* __restrict__ sorted_pad, const int32_t* __restrict__ step_block, const int64_t* __restrict__ step0,
                      int64_t n_rows, __nv_bfloat16* __restrict__ out) {
    constexpr int R = (NQ <= 2) ? 8 : (NQ <= 4 ? 4 : 1);       // rows per warp iteration (register budget)
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nq = D >> 2, nqp = Dp >> 2;
    const int64_t n_runs = n_rows / R;                          // n_rows is a multiple of 256
    for (int64_t run = warp0; run < n_runs; run += nwarps) {
        const int64_t row0 = run * R;
        const int64_t step = row0 / PA_NB;
        const int j0 = (int)(row0 - step * PA_NB);
        const int b = step_block[step];
        const int64_t t = step - step0[b];
        int64_t my_src = -1;
        if (lane < R) {
            const int g = sorted_pad[b * PA_NB + j0 + lane];
            if (g >= 0) {
                const int64_t s0 = goff[g];
                if (t < goff[g + 1] - s0) my_src = s0 + t;
            }
        }
        float4 v[R][NQ];
        int64_t src[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            src[r] = __shfl_sync(0xffffffffu, my_src, r);
            const float4* xr = reinterpret_cast<const float4*>(x + (src[r] < 0 ? 0 : src[r]) * (int64_t)D);
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                int q = lane + 32 * i;
                v[r][i] = (src[r] >= 0 && q < nq) ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                double a = (double)v[r][i].x, bb = (double)v[r][i].y, c = (double)v[r][i].z, d = (double)v[r][i].w;
                s = fma(a, a, s);
                s = fma(bb, bb, s);
                s = fma(c, c, s);
                s = fma(d, d, s);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
            float nrm = (float)sqrt(s);
            float den = nrm > 1e-12f ? nrm : 1e-12f;
            const float inv = __fdiv_rn(1.0f, den);
            uint2* orow = reinterpret_cast<uint2*>(out + (row0 + r) * (int64_t)Dp);
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                int q = lane + 32 * i;
                if (q < nqp) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(__fmul_rn(v[r][i].x, inv), __fmul_rn(v[r][i].y, inv));
                    __nv_bfloat162 hi = __floats2bfloat162_rn(__fmul_rn(v[r][i].z, inv), __fmul_rn(v[r][i].w, inv));
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    orow[q] = pk;                               // zero rows: v == 0 -> 0 * inv == 0
                }
            }
            for (int q = lane + 32 * NQ; q < nqp; q += 32) orow[q] = make_uint2(0u, 0u);
        }
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------------
struct PaParams {
    PgParams pg;                   // slot arrays, P, tau, RB, g_base (= first SORTED POSITION of this batch)
    const int32_t* sorted_pad;     // [n_blocks*256] label group of every block slot, -1 = padding
    const int32_t* blockT;         // [n_blocks] steps of each block
    const int64_t* step0;          // [n_blocks+1] first step of each block
    int32_t block_lo, n_blocks;    // blocks [block_lo, block_lo + n_blocks) of this batch
};

template <int KCH, int MT, int STAGES>
__global__ void __launch_bounds__(MT == 2 ? PG_THREADS2 : PG_THREADS, 1)
k_poolacc(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const PaParams q) {
    constexpr int NC = PA_NB;
    constexpr uint32_t A_TILE = 128 * 128;
    constexpr uint32_t A_BYTES = MT * KCH * A_TILE;
    constexpr uint32_t B_STAGE = NC * 128;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    static_assert(MT * NC <= 512, "TMEM columns");
    const PgParams& p = q.pg;

    extern __shared__ uint8_t pg_smem_raw[];
    const uint32_t raw = pg_smem_u32(pg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + A_BYTES;
    const uint32_t sBar = sB + STAGES * B_STAGE;
    const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8;
    const uint32_t bar_b_full = sBar + 16, bar_b_empty = bar_b_full + 8 * STAGES;
    const uint32_t bar_t_full = bar_b_empty + 8 * STAGES, bar_t_empty = bar_t_full + 8;
    const uint32_t s_tmem = bar_t_empty + 8;
    uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(pg_smem_raw + (s_tmem - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        pg_mbar_init(bar_a_full, 1);
        pg_mbar_init(bar_a_empty, 1);
        for (int s = 0; s < STAGES; ++s) { pg_mbar_init(bar_b_full + 8 * s, 1); pg_mbar_init(bar_b_empty + 8 * s, 1); }
        pg_mbar_init(bar_t_full, 1);
        pg_mbar_init(bar_t_empty, 4 * MT);
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

    const int64_t n_units = (int64_t)q.n_blocks * p.RB;        // unit = (block of 256 groups, row block of MT*128 bank rows)

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_phase = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int32_t bl = (int32_t)(u / p.RB), rb = (int32_t)(u - (int64_t)bl * p.RB);
                const int32_t b = q.block_lo + bl;
                const int32_t T = q.blockT[b];
                if (T <= 0) continue;
                const int64_t s0 = q.step0[b];
                pg_mbar_wait(bar_a_empty, a_phase ^ 1);
                pg_mbar_expect_tx(bar_a_full, A_BYTES);
#pragma unroll 1
                for (int rt = 0; rt < MT; ++rt)
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc)
                        pg_tma_load_2d(sA + (rt * KCH + kc) * A_TILE, &tmapA, kc * 64, (int32_t)((int64_t)rb * MT * 128 + rt * 128), bar_a_full);
                a_phase ^= 1;
                for (int32_t t = 0; t < T; ++t) {
                    const int32_t crow = (int32_t)((s0 + t) * NC);
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
        // ================= MMA issuer: all steps of the block accumulate into the same MT tiles =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_phase = 0, uidx = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int32_t bl = (int32_t)(u / p.RB);
                const int32_t T = q.blockT[q.block_lo + bl];
                if (T <= 0) continue;
                pg_mbar_wait(bar_a_full, a_phase);
                a_phase ^= 1;
                pg_mbar_wait(bar_t_empty, (uidx & 1u) ^ 1u);           // previous unit's tiles have been read out
                pg_fence_after();
                for (int32_t t = 0; t < T; ++t) {
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc) {
                        pg_mbar_wait(bar_b_full + 8 * stage, phase);
                        pg_fence_after();
                        const uint64_t db = pg_make_desc(sB + stage * B_STAGE);
#pragma unroll
                        for (int rt = 0; rt < MT; ++rt) {
                            const uint64_t da = pg_make_desc(sA + (rt * KCH + kc) * A_TILE);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                pg_mma_bf16(tmem_base + rt * NC, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC,
                                            (t | kc | kk) != 0 ? 1u : 0u);
                        }
                        pg_commit(bar_b_empty + 8 * stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                pg_commit(bar_t_full);
                pg_commit(bar_a_empty);
                ++uidx;
            }
        }
    } else {
        // ================= epilogue: warpgroup wg <-> row tile; thread <-> bank row; column <-> label group ==========
        const int wq = warp & 3;
        const int wg = (warp - 2) >> 2;
        const int rt = MT == 2 ? wg : 0;
        const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
        uint32_t uidx = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int32_t bl = (int32_t)(u / p.RB), rb = (int32_t)(u - (int64_t)bl * p.RB);
            const int32_t b = q.block_lo + bl;
            if (q.blockT[b] <= 0) continue;
            const int64_t tile128 = (int64_t)rb * MT + rt;
            const int64_t row = tile128 * 128 + wq * 32 + lane;
            const int64_t nsub = (int64_t)p.RB * MT * 4;
            // lane l holds the group ids / sizes of columns l, l+32, ..: 8 coalesced loads instead of 256 serial ones
            int32_t gid[NC / 32];
            float ginv[NC / 32];
#pragma unroll
            for (int i = 0; i < NC / 32; ++i) {
                const int g = q.sorted_pad[b * NC + i * 32 + lane];
                gid[i] = g;
                const int64_t n = g >= 0 ? p.goff[g + 1] - p.goff[g] : 0;
                ginv[i] = n > 0 ? 1.0f / (float)n : 0.f;
            }
            pg_mbar_wait(bar_t_full, uidx & 1u);
            pg_fence_after();
#pragma unroll
            for (int blk = 0; blk < NC / 32; ++blk) {
                float v[32];
                pg_tmem_ld32(tmem_base + lane_base + rt * NC + blk * 32, v);
                pg_tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    const int g = __shfl_sync(0xffffffffu, gid[blk], cc);
                    const float inv = __shfl_sync(0xffffffffu, ginv[blk], cc);
                    if (g < 0 || inv == 0.f) continue;                                // padding slot or empty group (uniform)
                    const float val = v[cc] * inv;
                    const bool pass = (val >= p.tau) && (row < p.P);
                    const uint32_t mpass = __ballot_sync(0xffffffffu, pass);
                    if (mpass == 0) continue;                                         // slot counts were zeroed before the launch
                    const int64_t pos = (int64_t)b * NC + blk * 32 + cc - p.g_base;   // sorted position inside the batch
                    pg_flush_write_call(&p, val, pass, mpass, pos * nsub + tile128 * 4 + wq, lane, row);
                }
            }
            pg_fence_before();
            __syncwarp();
            if (lane == 0) pg_mbar_arrive(bar_t_empty);
            ++uidx;
        }
    }

    pg_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------
template <int KCH, int MT, int STAGES>
static int pa_launch_t(sdk_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const PaParams& q, int grid) {
    constexpr size_t smem = (size_t)MT * KCH * 16384 + (size_t)STAGES * PA_NB * 128 + 256 + 1024;
    static_assert(smem <= PG_SMEM_LIMIT, "shared memory budget");
    auto kern = k_poolacc<KCH, MT, STAGES>;
    SDK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, MT == 2 ? PG_THREADS2 : PG_THREADS, smem, c->stream>>>(ta, tb, q);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

static int pa_mt_for(int kch) { return kch <= 4 ? 2 : 1; }

static int pa_launch(sdk_ctx* c, int kch, const CUtensorMap& ta, const CUtensorMap& tb, const PaParams& q, int grid) {
    switch (kch) {
        case 1: return pa_launch_t<1, 2, 6>(c, ta, tb, q, grid);
        case 2: return pa_launch_t<2, 2, 5>(c, ta, tb, q, grid);
        case 3: return pa_launch_t<3, 2, 4>(c, ta, tb, q, grid);
        case 4: return pa_launch_t<4, 2, 3>(c, ta, tb, q, grid);
        case 5: return pa_launch_t<5, 1, 4>(c, ta, tb, q, grid);
        case 6: return pa_launch_t<6, 1, 4>(c, ta, tb, q, grid);
        case 7: return pa_launch_t<7, 1, 3>(c, ta, tb, q, grid);
        default: return pa_launch_t<8, 1, 3>(c, ta, tb, q, grid);
    }
}

int sdk_poolacc_applicable(int32_t Dp, int32_t G, int32_t pool) {
    return sdk_poolgemm_supported(Dp) && pool == SDK_POOL_MEAN && G >= 128;
}

// Plan of the interleaved layout: groups sorted by size (counting sort), blocks of 256 groups, steps per block.
// Returns the total number of 256-row steps in *steps_out (one stream sync); S*256 / N - 1 is the zero padding.
int sdk_poolacc_plan(sdk_ctx* c, const int64_t* d_goff, int32_t G, int64_t* steps_out) {
    const int32_t n_blocks = (G + PA_NB - 1) / PA_NB;
    const int32_t Gpad = n_blocks * PA_NB;
    // ---- plan ----
    SDK_TRY(sdk_reserve(c, c->pa_hist, (size_t)PA_BUCKETS * 4));
    SDK_TRY(sdk_reserve(c, c->pa_sorted, (size_t)G * 4));
    SDK_TRY(sdk_reserve(c, c->pa_pos, (size_t)G * 4));
    SDK_TRY(sdk_reserve(c, c->pa_sorted_pad, (size_t)Gpad * 4));
    SDK_TRY(sdk_reserve(c, c->pa_blockT, (size_t)n_blocks * 4));
    SDK_TRY(sdk_reserve(c, c->pa_step0, (size_t)(n_blocks + 2) * 8));
    SDK_TRY(sdk_reserve(c, c->pa_seg_base, (size_t)G * 8));
    int32_t* hist = (int32_t*)c->pa_hist.p;
    int64_t* step0 = (int64_t*)c->pa_step0.p;
    int64_t* plan_total = step0 + n_blocks + 1;
    {
        sdk_prof_scope ps(c, "plan");
        SDK_CUDA(c, cudaMemsetAsync(hist, 0, (size_t)PA_BUCKETS * 4, c->stream));
        k_pa_hist<<<(G + 255) / 256, 256, 0, c->stream>>>(d_goff, G, hist);
        k_pa_scan<<<1, 1024, 0, c->stream>>>(hist);
        k_pa_scatter<<<(G + 255) / 256, 256, 0, c->stream>>>(d_goff, G, hist, (int32_t*)c->pa_sorted.p, (int32_t*)c->pa_pos.p);
        k_pa_blocks<<<(n_blocks + 127) / 128, 128, 0, c->stream>>>(d_goff, (const int32_t*)c->pa_sorted.p, G, n_blocks, Gpad,
                                                                   (int32_t*)c->pa_sorted_pad.p, (int32_t*)c->pa_blockT.p, step0, plan_total);
        k_pa_steps<<<1, 32, 0, c->stream>>>((const int32_t*)c->pa_blockT.p, n_blocks, step0, plan_total);
        k_pa_group_rows<<<(G + 255) / 256, 256, 0, c->stream>>>((const int32_t*)c->pa_pos.p, step0, G, (int64_t*)c->pa_seg_base.p);
        c->launches += 6;
        SDK_CUDA(c, cudaGetLastError());
    }
    int64_t S = 0;
    SDK_CUDA(c, cudaMemcpyAsync(&S, plan_total, 8, cudaMemcpyDeviceToHost, c->stream));
    SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    *steps_out = S;
    return SDK_OK;
}

// Normalises the raw segments into the planned interleaved layout, runs the accumulate-pooling GEMM and the slot
// merge.  Outputs: candidate rows + bound per label group, and (seg_base, stride) of every group's segments in the
// interleaved bf16 matrix c->seg_bf16 for the canonical re-score.  `S` = steps from sdk_poolacc_plan.
int sdk_launch_poolacc_candidates(sdk_ctx* c, const float* d_seg_raw, int64_t N, int32_t D, int32_t Dp, const __nv_bfloat16* d_bank,
                                  int64_t P, const int64_t* d_goff, int32_t G, int64_t S, float tau, int32_t ncand,
                                  int32_t* d_cand_row, float* d_gbound, const int64_t** d_seg_base_out, int64_t* seg_stride_out) {
    if (!sdk_poolacc_applicable(Dp, G, SDK_POOL_MEAN) || !c->tmap_encode) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path unavailable");
    if (P > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "tcgen05 path: at most 2^31-1 bank rows");
    if (D % 4 != 0 || D > 2048) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path needs D % 4 == 0");
    const int kch = Dp / 64, MT = pa_mt_for(kch);
    const int32_t n_blocks = (G + PA_NB - 1) / PA_NB;
    int64_t* step0 = (int64_t*)c->pa_step0.p;
    const int64_t n_rows = S * PA_NB;
    if (n_rows > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path: interleaved matrix exceeds 2^31-1 rows");
    *d_seg_base_out = (const int64_t*)c->pa_seg_base.p;
    *seg_stride_out = PA_NB;
    if (S == 0) {
        SDK_CUDA(c, cudaMemsetAsync(d_cand_row, 0xff, (size_t)G * ncand * 4, c->stream));
        return SDK_OK;
    }
    SDK_TRY(sdk_reserve(c, c->pa_step_block, (size_t)S * 4));
    SDK_TRY(sdk_reserve(c, c->seg_bf16, (size_t)n_rows * Dp * 2));
    k_pa_step_table<<<(unsigned)((S + 255) / 256), 256, 0, c->stream>>>(step0, n_blocks, S, (int32_t*)c->pa_step_block.p);
    c->launches++;
    // ---- K1, gather form ----
    {
        sdk_prof_scope ps(c, "normalize");
        int64_t blocks64 = (n_rows / 8 + 7) / 8;                    // a warp handles runs of up to 8 rows
        if (blocks64 < 1) blocks64 = 1;
        int blocks = (int)(blocks64 < (int64_t)c->sm_count * 8 ? blocks64 : (int64_t)c->sm_count * 8);
        int nq = (D / 4 + 31) / 32;
#define PA_NORM(NQ) k_pa_normalize_gather<NQ><<<blocks, 256, 0, c->stream>>>(d_seg_raw, D, Dp, d_goff, (const int32_t*)c->pa_sorted_pad.p, \
        (const int32_t*)c->pa_step_block.p, step0, n_rows, (__nv_bfloat16*)c->seg_bf16.p)
        if (nq <= 1) PA_NORM(1);
        else if (nq <= 2) PA_NORM(2);
        else if (nq <= 4) PA_NORM(4);
        else if (nq <= 8) PA_NORM(8);
        else PA_NORM(16);
#undef PA_NORM
        c->launches++;
        SDK_CUDA(c, cudaGetLastError());
    }
    // ---- GEMM + merge, in batches of blocks that keep the candidate slots under ~6 GB ----
    const int64_t rows_per_block = (int64_t)MT * 128;
    const int32_t RB = (int32_t)((P + rows_per_block - 1) / rows_per_block);
    const int32_t nsub = RB * MT * 4;
    const size_t per_group = (size_t)nsub * (PG_CS * 8 + 8);
    int64_t bbatch = (int64_t)((size_t)(6144ull << 20) / (per_group * PA_NB));
    if (bbatch < 1) bbatch = 1;
    if (bbatch > n_blocks) bbatch = n_blocks;
    SDK_TRY(sdk_reserve(c, c->slot_cnt, (size_t)bbatch * PA_NB * nsub * 4));
    SDK_TRY(sdk_reserve(c, c->slot_bound, (size_t)bbatch * PA_NB * nsub * 4));
    SDK_TRY(sdk_reserve(c, c->slot_row, (size_t)bbatch * PA_NB * nsub * PG_CS * 4));
    SDK_TRY(sdk_reserve(c, c->slot_val, (size_t)bbatch * PA_NB * nsub * PG_CS * 4));
    CUtensorMap ta, tb;
    SDK_TRY(pg_make_tmap(c, &ta, d_bank, P, Dp, 128));
    SDK_TRY(pg_make_tmap(c, &tb, c->seg_bf16.p, n_rows, Dp, PA_NB));
    (void)N;
    for (int64_t ba = 0; ba < n_blocks; ba += bbatch) {
        const int64_t bb = std::min<int64_t>(n_blocks, ba + bbatch);
        PaParams q;
        q.pg.goff = d_goff;
        q.pg.range_g = nullptr;
        q.pg.n_ranges = 0;
        q.pg.RB = RB;
        q.pg.P = P;
        q.pg.g_base = (int32_t)(ba * PA_NB);
        q.pg.pool = SDK_POOL_MEAN;
        q.pg.tau = tau;
        q.pg.mode = 0;
        q.pg.slot_cnt = (int32_t*)c->slot_cnt.p;
        q.pg.slot_row = (int32_t*)c->slot_row.p;
        q.pg.slot_val = (float*)c->slot_val.p;
        q.pg.slot_bound = (float*)c->slot_bound.p;
        q.pg.dense_out = nullptr;
        q.pg.dense_ld = 0;
        q.sorted_pad = (const int32_t*)c->pa_sorted_pad.p;
        q.blockT = (const int32_t*)c->pa_blockT.p;
        q.step0 = step0;
        q.block_lo = (int32_t)ba;
        q.n_blocks = (int32_t)(bb - ba);
        const int64_t n_units = (bb - ba) * RB;
        const int grid = (int)std::min<int64_t>(n_units, c->sm_count);
        // slots of groups that are never flushed (padding / empty) must read as empty
        SDK_CUDA(c, cudaMemsetAsync(c->slot_cnt.p, 0, (size_t)(bb - ba) * PA_NB * nsub * 4, c->stream));
        {
            sdk_prof_scope ps(c, "poolgemm");
            SDK_TRY(pa_launch(c, kch, ta, tb, q, grid));
        }
        pg_launch_merge(c, d_goff, (int32_t)(ba * PA_NB), (int32_t)((bb - ba) * PA_NB), nsub, (const int32_t*)c->pa_sorted_pad.p, tau, ncand,
                        d_cand_row, d_gbound);
        SDK_CUDA(c, cudaGetLastError());
    }
    return SDK_OK;
}
