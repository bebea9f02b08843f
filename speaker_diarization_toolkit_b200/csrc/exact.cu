// exact.cu -- K2a: canonical (bit-defined) pooled scores on the CUDA cores.
//
// This is (a) the "small bank" path of north_star (launch-latency-bound shapes: configs 1 and 2),
// (b) stage B of the tcgen05 path: re-scoring of the few candidate rows per label in the canonical
// arithmetic, and (c) the exhaustive fallback for label groups whose top-k certificate failed.
//
// Arithmetic = oracle/canonical.c steps (3)-(4): per pair an fp64 fma chain over ascending d
// (products of fp32/bf16 operands are exact in fp64), q = rint(score * 2^30), then integer
// sum / max per label group.  Integer pooling is order independent, so warps and CTAs combine their
// partial pools with shuffles and 64-bit atomics and the result is still bit-defined.
//
// Mapping: CTA = (label group, tile of RT bank-row slots, segment split z); thread = one segment,
// RT fp64 accumulators.  The RT bank rows of the tile are converted to fp64 once per 128-d chunk
// into shared memory and read back as warp-wide broadcasts (no bank conflicts, no per-thread
// conversion); each thread streams its own segment row with 128-bit loads.
#include "common.cuh"

#define SDK_EX_THREADS 256
#define SDK_EX_DC 128

__device__ __forceinline__ long long sdk_warp_sum_ll(long long v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ long long sdk_warp_max_ll(long long v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        long long o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o > v ? o : v;
    }
    return v;
}

// acc[r] += sum_d x[d] * bs[d][r] for r < A (fp64 fma chain over ascending d: the canonical order)
template <int A, int RT, bool BF16>
__device__ __forceinline__ void sdk_ex_dot(double (&acc)[RT], const double (&bs)[SDK_EX_DC][RT], const void* __restrict__ seg_ops,
                                           int64_t srow, int32_t pitch, int d0, int dc) {
    if (BF16) {
        const __nv_bfloat16* xr = reinterpret_cast<const __nv_bfloat16*>(seg_ops) + srow * (int64_t)pitch + d0;
        int dd = 0;
        for (; dd + 8 <= dc; dd += 8) {      // pitch % 8 == 0 and d0 % 8 == 0 -> 16-byte aligned
            uint4 pk = __ldg(reinterpret_cast<const uint4*>(xr + dd));
            uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                double x0 = (double)__uint_as_float(w[h] << 16);
                double x1 = (double)__uint_as_float(w[h] & 0xffff0000u);
#pragma unroll
                for (int r = 0; r < A; ++r) acc[r] = fma(x0, bs[dd + 2 * h][r], acc[r]);
#pragma unroll
                for (int r = 0; r < A; ++r) acc[r] = fma(x1, bs[dd + 2 * h + 1][r], acc[r]);
            }
        }
        for (; dd < dc; ++dd) {
            double x = (double)__bfloat162float(xr[dd]);
#pragma unroll
            for (int r = 0; r < A; ++r) acc[r] = fma(x, bs[dd][r], acc[r]);
        }
    } else {
        const float* xr = reinterpret_cast<const float*>(seg_ops) + srow * (int64_t)pitch + d0;
        int dd = 0;
        if ((pitch & 3) == 0) {
            for (; dd + 4 <= dc; dd += 4) {
                float4 pk = __ldg(reinterpret_cast<const float4*>(xr + dd));
                float w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    double x = (double)w[h];
#pragma unroll
                    for (int r = 0; r < A; ++r) acc[r] = fma(x, bs[dd + h][r], acc[r]);
                }
            }
        }
        for (; dd < dc; ++dd) {
            double x = (double)__ldg(xr + dd);
#pragma unroll
            for (int r = 0; r < A; ++r) acc[r] = fma(x, bs[dd][r], acc[r]);
        }
    }
}

template <int RT, bool BF16>
__global__ void __launch_bounds__(SDK_EX_THREADS)
k_exact_q30(const void* __restrict__ seg_ops, const void* __restrict__ bank_ops, int32_t D, int32_t pitch,
            const int64_t* __restrict__ goff, const int32_t* __restrict__ glist,
            const int32_t* __restrict__ cand_row, int64_t nslot, int32_t ntiles, int32_t pool,
            long long* __restrict__ qpool, const PaGroup* __restrict__ grp, int64_t n_seg) {
    __shared__ double bs[SDK_EX_DC][RT];
    __shared__ int32_t srow[RT];
    const int tid = threadIdx.x;
    const int32_t gi = blockIdx.x / ntiles;
    const int32_t tile = blockIdx.x - gi * ntiles;
    const int32_t g = glist ? glist[gi] : gi;
    int64_t s0 = goff[g], s1 = goff[g + 1];
    if (n_seg >= 0) {       // labels not validated yet (exact path: the flag is read at fetch time): never read past the input
        s0 = s0 < 0 ? 0 : (s0 > n_seg ? n_seg : s0);
        s1 = s1 < 0 ? 0 : (s1 > n_seg ? n_seg : s1);
    }
    if (s1 <= s0) return;
    // row of segment t in seg_ops: rbase + (t / gc) * gstep + t % gc   (plain layout: gc == 1, gstep == 1)
    int64_t rbase = s0, gstep = 1;
    int32_t gc = 1;
    if (grp) { const PaGroup pg = grp[g]; rbase = pg.base; gc = pg.c; gstep = 256; }
    if (tid < RT) {
        int64_t slot = (int64_t)tile * RT + tid;
        int32_t r = -1;
        if (slot < nslot) r = cand_row ? cand_row[(int64_t)gi * nslot + slot] : (int32_t)slot;
        srow[tid] = r;
    }
    __syncthreads();
    int nact = 0;                       // 1: only slot 0 holds a row (front-packed candidate list); else treat all RT
#pragma unroll
    for (int r = 0; r < RT; ++r) nact += srow[r] >= 0 ? 1 : 0;
    if (nact == 0) return;
    if (!(nact == 1 && srow[0] >= 0)) nact = RT;

    const int nz = gridDim.y;
    for (int64_t cbase = s0 + (int64_t)blockIdx.y * SDK_EX_THREADS; cbase < s1; cbase += (int64_t)nz * SDK_EX_THREADS) {
        const int64_t s = cbase + tid;
        const bool valid = s < s1;
        const int32_t t = (int32_t)(s - s0);
        const int64_t srow_of = gc == 1 ? rbase + (int64_t)t * gstep : rbase + (int64_t)(t / gc) * gstep + (t % gc);
        double acc[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) acc[r] = 0.0;
        for (int d0 = 0; d0 < D; d0 += SDK_EX_DC) {
            const int dc = min(SDK_EX_DC, D - d0);
            __syncthreads();
            for (int idx = tid; idx < dc * RT; idx += SDK_EX_THREADS) {
                int r = idx / dc, dd = idx - r * dc;     // consecutive threads -> consecutive d (coalesced)
                int32_t row = srow[r];
                double v = 0.0;
                if (row >= 0) {
                    if (BF16) v = (double)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bank_ops)[(int64_t)row * pitch + d0 + dd]);
                    else v = (double)reinterpret_cast<const float*>(bank_ops)[(int64_t)row * pitch + d0 + dd];
                }
                bs[dd][r] = v;
            }
            __syncthreads();
            if (valid) {
                // the candidate lists of the sparse path are front-packed and usually hold ONE row: do not spend
                // RT fp64 fma chains on the empty slots
                if (nact == 1) sdk_ex_dot<1, RT, BF16>(acc, bs, seg_ops, srow_of, pitch, d0, dc);
                else sdk_ex_dot<RT, RT, BF16>(acc, bs, seg_ops, srow_of, pitch, d0, dc);
            }
        }
        // fixed point + integer pooling over this CTA's segments
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            long long q;
            if (pool == 0) {
                q = valid ? __double2ll_rn(acc[r] * SDK_Q30) : 0ll;
                q = sdk_warp_sum_ll(q);
            } else {
                q = valid ? __double2ll_rn(acc[r] * SDK_Q30) : LLONG_MIN;
                q = sdk_warp_max_ll(q);
            }
            if ((tid & 31) == 0 && srow[r] >= 0) {
                long long* dst = qpool + (int64_t)gi * nslot + (int64_t)tile * RT + r;
                if (pool == 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)q);
                else atomicMax(dst, q);
            }
        }
    }
}

// ---- dense scans (every bank row against every segment of a label): a register-tiled fp64 "GEMM" -----------------------
// The thread-per-segment kernel above re-reads a label's segments once per row tile and feeds every fp64 fma from its own
// shared-memory broadcast; fine for a handful of candidate rows, wasteful for a whole bank.  Here a CTA owns a tile of
// 64 segments x 64 bank rows of one label: both operands are staged once per 32-wide d chunk as fp64 in shared memory
// (coalesced global loads), a thread keeps a 4 x 4 block of pair accumulators in registers (two 128-bit broadcasts +
// two 128-bit loads feed 16 fmas).  The arithmetic per pair is unchanged -- one fp64 fma chain over ascending d, then
// q = rint(score * 2^30) -- and the integer pooling over segments goes through shared-memory and global 64-bit atomics.
#define SDK_EXD_TS 64          // segments per tile
#define SDK_EXD_TR 64          // bank rows per tile
#define SDK_EXD_DC 32          // d chunk
#define SDK_EXD_LD 66          // padded leading dimension of the staged tiles (doubles)

template <bool BF16>
__device__ __forceinline__ float sdk_exd_load(const void* __restrict__ ops, int64_t row, int32_t pitch, int d) {
    if (BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ops)[row * (int64_t)pitch + d]);
    return reinterpret_cast<const float*>(ops)[row * (int64_t)pitch + d];
}

template <bool BF16, int RPT /* bank rows per thread: 4 -> 64-row tiles, 2 -> 32-row tiles (more CTAs for small banks) */>
__global__ void __launch_bounds__(256)
k_exact_dense(const void* __restrict__ seg_ops, const void* __restrict__ bank_ops, int32_t D, int32_t pitch,
              const int64_t* __restrict__ goff, const int32_t* __restrict__ glist, int64_t P, int32_t ntiles, int32_t pool,
              long long* __restrict__ qpool, const PaGroup* __restrict__ grp, int64_t n_seg) {
    __shared__ __align__(16) double s_seg[SDK_EXD_DC][SDK_EXD_LD];
    __shared__ __align__(16) double s_row[SDK_EXD_DC][SDK_EXD_LD];
    constexpr int TR = 16 * RPT;                                      // bank rows per tile
    __shared__ long long s_pool[SDK_EXD_TR];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;        // rows RPT*tx.., segments 4*ty..
    const int32_t gi = blockIdx.x / ntiles;
    const int32_t tile = blockIdx.x - gi * ntiles;
    const int32_t g = glist ? glist[gi] : gi;
    int64_t s0 = goff[g], s1 = goff[g + 1];
    if (n_seg >= 0) {
        s0 = s0 < 0 ? 0 : (s0 > n_seg ? n_seg : s0);
        s1 = s1 < 0 ? 0 : (s1 > n_seg ? n_seg : s1);
    }
    if (s1 <= s0) return;
    int64_t rbase = s0, gstep = 1;
    int32_t gc = 1;
    if (grp) { const PaGroup pg = grp[g]; rbase = pg.base; gc = pg.c; gstep = 256; }
    const int64_t row0 = (int64_t)tile * TR;
    const int nz = gridDim.y;
    for (int64_t cbase = s0 + (int64_t)blockIdx.y * SDK_EXD_TS; cbase < s1; cbase += (int64_t)nz * SDK_EXD_TS) {
        double acc[4][RPT];                                           // [segment][row]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < RPT; ++j) acc[i][j] = 0.0;
        if (tid < TR) s_pool[tid] = pool == 0 ? 0ll : LLONG_MIN;
        // staging: thread -> (item = tid / 32 + 8 * it, d = tid % 32): consecutive lanes read consecutive d (coalesced).  The
        // global loads of chunk c+1 are issued before the fmas of chunk c and land in registers (software pipeline): a
        // d chunk is only 32 x 16 fmas per thread, far too short to hide an L2 round trip otherwise.
        const int dd = tid & 31;
        int64_t seg_row[8];
        bool seg_live[8], row_live[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int item = (tid >> 5) + 8 * it;
            const int64_t s = cbase + item;
            seg_live[it] = s < s1;
            const int32_t t = (int32_t)(s - s0);
            seg_row[it] = gc == 1 ? rbase + (int64_t)t * gstep : rbase + (int64_t)(t / gc) * gstep + (t % gc);
            row_live[it] = item < TR && row0 + item < P;
        }
        float pre_s[8], pre_r[8];                                     // raw operands (fp32 holds bf16 exactly); widened at the store
        auto fetch = [&](int d0) {
            const bool dlive = d0 + dd < D;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int item = (tid >> 5) + 8 * it;
                pre_s[it] = (dlive && seg_live[it]) ? sdk_exd_load<BF16>(seg_ops, seg_row[it], pitch, d0 + dd) : 0.f;
                pre_r[it] = (dlive && row_live[it]) ? sdk_exd_load<BF16>(bank_ops, row0 + item, pitch, d0 + dd) : 0.f;
            }
        };
        fetch(0);
        for (int d0 = 0; d0 < D; d0 += SDK_EXD_DC) {
            __syncthreads();                                          // everyone is done reading the previous chunk
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int item = (tid >> 5) + 8 * it;
                s_seg[dd][item] = (double)pre_s[it];
                s_row[dd][item] = (double)pre_r[it];
            }
            __syncthreads();
            if (d0 + SDK_EXD_DC < D) fetch(d0 + SDK_EXD_DC);          // in flight while this chunk is contracted
#pragma unroll 8
            for (int k = 0; k < SDK_EXD_DC; ++k) {
                const double2 a01 = *reinterpret_cast<const double2*>(&s_seg[k][4 * ty]);
                const double2 a23 = *reinterpret_cast<const double2*>(&s_seg[k][4 * ty + 2]);
                const double a[4] = {a01.x, a01.y, a23.x, a23.y};
                double b[RPT];
#pragma unroll
                for (int j = 0; j < RPT; j += 2) {
                    const double2 bb = *reinterpret_cast<const double2*>(&s_row[k][RPT * tx + j]);
                    b[j] = bb.x;
                    b[j + 1] = bb.y;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < RPT; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
        }
        // fixed point, pool this thread's 4 segments per row, then over the 16 threads that share the rows
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            long long q = pool == 0 ? 0ll : LLONG_MIN;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (cbase + 4 * ty + i < s1) {
                    const long long v = __double2ll_rn(acc[i][j] * SDK_Q30);
                    q = pool == 0 ? q + v : (v > q ? v : q);
                }
            }
            // lanes l and l ^ 16 hold the same rows (ty differs by one): combine before the shared-memory atomic
            const long long o = __shfl_xor_sync(0xffffffffu, q, 16);
            q = pool == 0 ? q + o : (o > q ? o : q);
            if ((tid & 16) == 0) {
                if (pool == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&s_pool[RPT * tx + j]), (unsigned long long)q);
                else atomicMax(&s_pool[RPT * tx + j], q);
            }
        }
        __syncthreads();
        if (tid < TR && row0 + tid < P) {
            long long* dst = qpool + (int64_t)gi * P + row0 + tid;
            if (pool == 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)s_pool[tid]);
            else atomicMax(dst, s_pool[tid]);
        }
    }
}

__global__ void k_fill_ll(long long* p, int64_t n, long long v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

int sdk_launch_exact(sdk_ctx* c, const void* d_seg_ops, const void* d_bank_ops, int32_t is_bf16, int32_t D,
                     int32_t pitch, const int64_t* d_goff, const int32_t* d_glist, int32_t ngroups,
                     const int32_t* d_cand_row, int64_t nslot, int32_t pool, long long* d_qpool, const PaGroup* d_grp, int64_t n_seg) {
    if (ngroups <= 0 || nslot <= 0) return SDK_OK;
    sdk_prof_scope ps(c, "exact");
    int64_t total = (int64_t)ngroups * nslot;
    int fb = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    k_fill_ll<<<fb, 256, 0, c->stream>>>(d_qpool, total, pool == 0 ? 0ll : LLONG_MIN);
    c->launches++;
    if (!d_cand_row && nslot >= SDK_EXD_TR) {
        // dense scan of a bank: register-tiled kernel; long labels are split over blockIdx.y (64 segments per pass)
        // 64-row tiles; 32-row tiles when that is what it takes to give every SM a couple of CTAs (one meeting vs a few
        // hundred profiles)
        int rpt = 4;
        if (((nslot + 63) / 64) * (int64_t)ngroups < 2 * (int64_t)c->sm_count) rpt = 2;
        const int64_t ntiles64 = (nslot + 16 * rpt - 1) / (16 * rpt);
        const int64_t blocks = ntiles64 * ngroups;
        if (blocks > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "exact path: too many (group,row-tile) blocks");
        int nz = 1;
        if (blocks < 4 * (int64_t)c->sm_count) {
            nz = (int)((4 * (int64_t)c->sm_count + blocks - 1) / blocks);
            if (nz > 64) nz = 64;
        }
        dim3 grid((unsigned)blocks, (unsigned)nz);
#define SDK_EXD_CASE(B, R) k_exact_dense<B, R><<<grid, 256, 0, c->stream>>>(d_seg_ops, d_bank_ops, D, pitch, d_goff, d_glist, nslot, (int)ntiles64, pool, d_qpool, d_grp, n_seg)
        if (is_bf16) { if (rpt == 4) SDK_EXD_CASE(true, 4); else SDK_EXD_CASE(true, 2); }
        else { if (rpt == 4) SDK_EXD_CASE(false, 4); else SDK_EXD_CASE(false, 2); }
#undef SDK_EXD_CASE
        c->launches++;
        SDK_CUDA(c, cudaGetLastError());
        return SDK_OK;
    }
    // row-slot tile width: wide tiles for dense scans, narrow for the few re-scored candidates
    // the candidate list is front-packed, so on the sparse path tiles of 4 slots let the empty tail exit at once
    int RT = d_cand_row ? 4 : (nslot >= 16 ? 16 : (nslot > 4 ? 8 : 4));
    // dense scans of a small bank (one meeting against a few hundred profiles): prefer enough CTAs for ~3 waves over
    // wide tiles
    while (!d_cand_row && RT > 4 && ((nslot + RT - 1) / RT) * (int64_t)ngroups < 3 * (int64_t)c->sm_count) RT /= 2;
    int64_t ntiles64 = (nslot + RT - 1) / RT;
    int64_t blocks = ntiles64 * ngroups;
    if (blocks > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "exact path: too many (group,row-tile) blocks");
    // split long groups over blockIdx.y when there are too few CTAs to fill the GPU
    int nz = 1;
    if (blocks < 2 * c->sm_count) {
        nz = (int)((2 * c->sm_count + blocks - 1) / blocks);
        if (nz > 64) nz = 64;
    }
    dim3 grid((unsigned)blocks, (unsigned)nz);
    int ntiles = (int)ntiles64;
#define SDK_EX_CASE(RTV)                                                                                   \
    do {                                                                                                   \
        if (is_bf16) k_exact_q30<RTV, true><<<grid, SDK_EX_THREADS, 0, c->stream>>>(d_seg_ops, d_bank_ops, D, pitch, d_goff, d_glist, d_cand_row, nslot, ntiles, pool, d_qpool, d_grp, n_seg); \
        else k_exact_q30<RTV, false><<<grid, SDK_EX_THREADS, 0, c->stream>>>(d_seg_ops, d_bank_ops, D, pitch, d_goff, d_glist, d_cand_row, nslot, ntiles, pool, d_qpool, d_grp, n_seg); \
    } while (0)
    if (RT == 16) SDK_EX_CASE(16);
    else if (RT == 8) SDK_EX_CASE(8);
    else SDK_EX_CASE(4);
#undef SDK_EX_CASE
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}
