// gemv.cu -- K2d: the small-query latency path.  A HANDFUL of query segments against a big bank ("8 pooled label centroids
// vs a 125k-row shard", BASELINE config 4 variant (i)): the whole identify call is TWO kernels and no host round trip.
//
//   k_gemv8       one HBM-bound pass over the bank (2*Dp bytes per bank row, nothing else), fused with everything that
//                 used to be separate launches in front of it: label-group offsets + label validation, the canonical
//                 normalise-and-cast of the <= 8 query rows (every CTA redoes it from the 16 KB of raw queries while its
//                 first bank tiles are already in flight; CTA 0 publishes the operand copies for stage B), and behind it:
//                 the per-CTA top-16 of every label's approximate pooled scores (+ a bound on everything dropped), so the
//                 kernel leaves 2 x grid short sorted lists per label instead of one candidate slot per 32 bank rows.
//   k_gemv8_tail  one CTA per label: merge of those lists (threshold from the list maxima, compaction, exact rank),
//                 canonical fp64 re-score of the candidates (stage B), row->speaker max, threshold, ordered top-k and the
//                 certificate -- what k_pg_merge + k_exact_q30 + k_select did in three launches over 4 MB of slots.
//                 Launched with programmatic stream serialisation: its launch latency hides under the bank stream.
// A failed certificate (rare) is not handled by a host check inside the call: the label is put on the fall-back list and
// the library settles it -- exhaustively, in the canonical arithmetic -- when the results are next fetched (api.cu).
//
// The stream: every CTA owns a contiguous range of the bank (nchunk * b / grid tiles of 32 rows, so all SMs stream the same
// amount +- one tile); tile j of a pass belongs to warp j % 16, which computes the complete dot products of its 32 rows with
// the 8 queries on mma.sync m16n8k16 (bf16 in, fp32 accumulate) -- bank rows are the M side, read STRAIGHT FROM GLOBAL
// MEMORY into the A fragments with 128-bit loads (see k_gemv8), the queries the N side (fragments pre-packed in shared
// memory).  Two TMA versions were measured first: per-warp rings of 4 KB K-chunk stages (round 1: 45 % of the HBM peak) and
// a CTA-wide ring of whole 32 KB row tiles (47 %; ncu: DRAM 38 % busy, every tile 8 us from request to release) -- with one
// elected producer and a handful of stages the number of bytes in flight per SM, not the DRAM, was the bound.
// Products of bf16 operands are exact in fp32; only the fp32 accumulation order differs from the canonical arithmetic,
// well inside the certificate's eps (tests/test_gpu_certificate.py measures it).
#include "tcgen05.cuh"

#define GV_NQ 8
#define GV_CW 16                     // consumer warps
#define GV_THREADS (32 * GV_CW)
#define GV_ROWS 32                   // bank rows per tile
#define GV_PASS_TILES 16             // tiles per pass (one per warp): their raw scores wait in shared memory for the selection,
                                     // which then runs while the next pass is already streaming in from DRAM
#define GV_PASS_ROWS (GV_ROWS * GV_PASS_TILES)
#define GV_RAW_LD (GV_PASS_ROWS + 8) // padded: the fragment stores of the four query pairs land in different banks
#define GV_KEEP 16                   // kept per (label, CTA, half of the pass)
#define GV_PARTS 2                   // selection warps per label: 8 labels x 2 = the 16 consumer warps
#define GV_TAIL_THREADS 256
#define GV_TAIL_CAP 1024             // compacted candidates per label in the tail

// Phase trace of the two kernels (tools/gv_trace.py builds a separate library with -DSDK_GV_TRACE; the shipped one has none
// of this): thread 0 of every CTA stamps clock64 at the phase boundaries, plus globaltimer at both ends.
#ifdef SDK_GV_TRACE
__device__ unsigned long long gv_trace[2][160][24];
__device__ __forceinline__ unsigned long long gv_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define GV_T(kern, slot) do { if (threadIdx.x == 0 && blockIdx.x < 160) gv_trace[kern][blockIdx.x][slot] = (unsigned long long)clock64(); } while (0)
#define GV_TG(kern, slot) do { if (threadIdx.x == 0 && blockIdx.x < 160) gv_trace[kern][blockIdx.x][slot] = gv_gtime(); } while (0)
extern "C" int sdk_debug_gv_trace(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, gv_trace, sizeof(gv_trace)) == cudaSuccess ? 0 : -1;
}
#else
#define GV_T(kern, slot) do { } while (0)
#define GV_TG(kern, slot) do { } while (0)
#endif

struct GvParams {
    const void* seg_raw;          // [N, D] raw query rows, fp32 or fp16
    const int32_t* seg_label;     // [N] global label ids
    int32_t in_dtype, label_base;
    int32_t N, L, D, Dp;
    int64_t P;
    int32_t pool;
    float tau;
    const __nv_bfloat16* bank;    // [P, Dp] normalised bf16 bank operands
    int32_t nchunk;               // 32-row tiles of the bank
    int32_t nslots;               // lists per query slot = grid * GV_PARTS
    // published by CTA 0 for the tail / the fall-back
    int64_t* goff;                // [L + 1]
    int32_t* flags;               // label validation bits
    __nv_bfloat16* seg_bf16;      // [N, Dp]
    float* seg_f32;               // [N, D] or null (fp32 banks)
    // lists: index (q0 * nslots + slot), q0 = first query of the label
    float* slot_bound;
    unsigned long long* slot_key; // [.., GV_KEEP] (score key << 32 | ~row), descending, 0 = no entry: every entry is written
};

// canonical normalise of one row (D <= 512) by one warp (oracle/canonical.c step (1); same element -> lane assignment and
// the same order as k_normalize_vec / k_normalize_generic): the bf16 copy goes STRAIGHT INTO THE B FRAGMENTS of query n in
// shared memory (layout: see k_gemv8), optional global copies.  All of the lane's elements are loaded BEFORE the first one
// is used (a load -> convert -> fma loop that waits for memory sixteen times in a row cost 6-8 us here, a third of the kernel).
template <typename TIn, typename AfterLoads>
__device__ __forceinline__ void gv_normalize_row(const TIn* __restrict__ xr, int32_t D, int32_t Dp, int lane, uint2* s_frag, int n,
                                                 __nv_bfloat16* g_bf16, float* g_f32, AfterLoads after_loads) {
    const int nq = D >> 2;
    const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(xr) % sdk_in<TIn>::align) == 0;
    float4 v[4];                                               // chunks lane, lane + 32, ..: elements 4q .. 4q + 3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = lane + 32 * i;
        if (vec) v[i] = q < nq ? sdk_in<TIn>::ld4(xr, q) : make_float4(0.f, 0.f, 0.f, 0.f);
        else {
            v[i].x = 4 * q + 0 < D ? sdk_in<TIn>::ld1(xr, 4 * q + 0) : 0.f;
            v[i].y = 4 * q + 1 < D ? sdk_in<TIn>::ld1(xr, 4 * q + 1) : 0.f;
            v[i].z = 4 * q + 2 < D ? sdk_in<TIn>::ld1(xr, 4 * q + 2) : 0.f;
            v[i].w = 4 * q + 3 < D ? sdk_in<TIn>::ld1(xr, 4 * q + 3) : 0.f;
        }
    }
    after_loads();                                             // (the bank prefetch goes out behind the query loads, not in front)
    GV_T(0, 18);
    double s = 0.0;                                            // (elements past D are zero: fma(0, 0, s) == s exactly)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double a = (double)v[i].x, b = (double)v[i].y, c = (double)v[i].z, d = (double)v[i].w;
        s = fma(a, a, s);
        s = fma(b, b, s);
        s = fma(c, c, s);
        s = fma(d, d, s);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
    const float nrm = (float)sqrt(s);
    const float den = nrm > 1e-12f ? nrm : 1e-12f;
    const float inv = __fdiv_rn(1.0f, den);
    GV_T(0, 19);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = lane + 32 * i;
        if (4 * q >= Dp) continue;
        const float o[4] = {__fmul_rn(v[i].x, inv), __fmul_rn(v[i].y, inv), __fmul_rn(v[i].z, inv), __fmul_rn(v[i].w, inv)};
        __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        // 32-bit words 2q, 2q + 1 of the row = k block q / 8, MMA j = q % 2, fragment lane (n, c = (q % 8) / 2)
        s_frag[((q >> 3) * 2 + (q & 1)) * 32 + n * 4 + ((q & 7) >> 1)] = pk;
        if (g_bf16) reinterpret_cast<uint2*>(g_bf16)[q] = pk;
        if (g_f32) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (4 * q + t < D) g_f32[4 * q + t] = o[t];
        }
    }
}

// 64-bit max over the warp (keys are unique or zero)
__device__ __forceinline__ unsigned long long gv_warp_max_u64(unsigned long long v) {
    const uint32_t hi = __reduce_max_sync(0xffffffffu, (uint32_t)(v >> 32));
    const uint32_t lo = __reduce_max_sync(0xffffffffu, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0u);
    return ((unsigned long long)hi << 32) | lo;
}

// Top-GV_KEEP of this warp's half of a pass (+ the list kept from earlier passes), by (score desc, row asc).
// vk[t] = orderable score key of row rbase + 32 t (0 = not a candidate).  The GV_KEEP-th largest of the 32 per-lane maxima
// is a lower bound T0 on the GV_KEEP-th best score, so only the handful of entries >= T0 are compacted (ballots) and ranked
// exactly, on full (score, row) keys; everything up to there works on the 32-bit score keys alone.  If more than 64 entries
// survive (clustered data), the extraction loop does it the slow way.  Returns the new kept entry of this lane (lanes
// 0 .. GV_KEEP-1, descending) and raises `bound` to the best dropped score.
// The phase is bound by instruction issue (all sixteen warps select at once, ~9 cycles per warp instruction: measured with
// tools/gv_trace.py), so what counts is the number of instructions: a bitonic network for T0, no per-tile bookkeeping.
__device__ __forceinline__ unsigned long long gv_select_top(const uint32_t (&vk)[GV_PASS_TILES / GV_PARTS], uint32_t rbase, unsigned long long kept,
                                                             float& bound, unsigned long long* s_cmp /*[64 + GV_KEEP + 2] per warp*/, int lane) {
    constexpr int T = GV_PASS_TILES / GV_PARTS;
    const uint32_t kv = (uint32_t)(kept >> 32);                // score key of the entry kept from earlier passes (0: none)
    uint32_t lmax = kv;
    int nl = kv != 0u ? 1 : 0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        lmax = vk[t] > lmax ? vk[t] : lmax;
        nl += vk[t] != 0u ? 1 : 0;
    }
    // the GV_KEEP-th largest lane maximum: bitonic sort of the 32 values across the lanes (descending), read at lane GV_KEEP-1
    uint32_t sv = lmax;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j >= 1; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, sv, j);
            const bool up = ((lane & k2) == 0) == ((lane & j) == 0);      // this lane keeps the larger of the pair
            sv = up ? (o > sv ? o : sv) : (o < sv ? o : sv);
        }
    }
    GV_T(0, 21);
    const uint32_t T0 = __shfl_sync(0xffffffffu, sv, GV_KEEP - 1);               // 0 when fewer than GV_KEEP lanes hold anything
    const uint32_t T0e = T0 != 0u ? T0 : 1u;                                     // (zero keys never pass)
    const int nlive = __reduce_add_sync(0xffffffffu, nl);
    // compaction of everything with a score >= T0
    int m = 0;
    const uint32_t lt = (1u << lane) - 1u;
    {
        const bool p = kv >= T0e;
        const uint32_t mk = __ballot_sync(0xffffffffu, p);
        if (p) s_cmp[__popc(mk & lt)] = kept;
        m = __popc(mk);
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const bool p = vk[t] >= T0e;
        const uint32_t mk = __ballot_sync(0xffffffffu, p);
        if (p) {
            const int pos = m + __popc(mk & lt);
            if (pos < 64) s_cmp[pos] = ((unsigned long long)vk[t] << 32) | (0xffffffffu - (rbase + 32u * (uint32_t)t));
        }
        m += __popc(mk);
    }
    __syncwarp();
    GV_T(0, 22);
#ifdef SDK_GV_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 160) gv_trace[0][blockIdx.x][12] = (unsigned long long)m;
#endif
    unsigned long long newkept = 0ull;
    if (m <= 64) {
        // exact rank of the survivors: lane holds entries lane and lane + 32
        const unsigned long long e0 = lane < m ? s_cmp[lane] : 0ull;
        int r0 = 0;
        unsigned long long e1 = 0ull;
        int r1 = 0;
        if (m <= 32) {
#pragma unroll 4
            for (int j = 0; j < m; ++j) r0 += s_cmp[j] > e0 ? 1 : 0;
        } else {
            e1 = lane + 32 < m ? s_cmp[lane + 32] : 0ull;
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
                const unsigned long long o = s_cmp[j];
                r0 += o > e0 ? 1 : 0;
                r1 += o > e1 ? 1 : 0;
            }
        }
        __syncwarp();
        // scatter by rank: positions 0 .. GV_KEEP-1 are the new list, position GV_KEEP is the best dropped entry
        if (e0 != 0ull && r0 <= GV_KEEP) s_cmp[64 + r0] = e0;
        if (e1 != 0ull && r1 <= GV_KEEP) s_cmp[64 + r1] = e1;
        __syncwarp();
        const int nk = m < GV_KEEP ? m : GV_KEEP;
        if (lane < nk) newkept = s_cmp[64 + lane];
        if (m > GV_KEEP) bound = fmaxf(bound, sdk_funkey((uint32_t)(s_cmp[64 + GV_KEEP] >> 32)));
        else if (nlive > m) bound = fmaxf(bound, sdk_funkey(T0));               // live entries below T0 were never compacted
        __syncwarp();
        return newkept;
    }
    // slow path: repeated extraction of the largest key below the last one
    unsigned long long last = ~0ull;
#pragma unroll 1
    for (int it = 0; it <= GV_KEEP; ++it) {
        unsigned long long best = (kept < last) ? kept : 0ull;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const unsigned long long c = vk[t] ? (((unsigned long long)vk[t] << 32) | (0xffffffffu - (rbase + 32u * (uint32_t)t))) : 0ull;
            best = (c < last && c > best) ? c : best;
        }
        best = gv_warp_max_u64(best);
        if (best == 0ull) break;
        if (it == GV_KEEP) { bound = fmaxf(bound, sdk_funkey((uint32_t)(best >> 32))); break; }
        if (lane == it) newkept = best;
        last = best;
    }
    return newkept;
}

// KB = Dp / 32 (K blocks of 32).  The bank is read straight from global memory into the A fragments: lane (g = lane / 4,
// c = lane % 4) loads 16 bytes -- the 8 bf16 at k = 32 kb + 8 c .. + 7 -- of bank rows g and g + 8 of a 16-row tile; the four
// lanes of a quad cover 64 contiguous bytes of a row, a warp instruction 8 rows x 64 B in full 32-byte sectors.  Those 8
// values feed TWO m16n8k16 MMAs (4 values each: the k index of a dot product may be permuted at will as long as both
// operands agree, so MMA j takes k = 32 kb + 8 c + 4 j + {0,1} for the fragment's columns {2c, 2c+1} and + {2,3} for
// {2c+8, 2c+9}); the query fragments are packed in shared memory in the same order.  16 independent 128-bit loads per lane
// are in flight before the first MMA of a group, sixteen warps per SM: plain loads with that much memory-level parallelism
// are what K1 reaches 82 % of the HBM peak with; the TMA ring version of this kernel stalled at 47 %.
template <int KB>
__global__ void __launch_bounds__(GV_THREADS, 1)
k_gemv8(const __grid_constant__ GvParams q) {
    extern __shared__ uint8_t gv_smem_raw[];
    // B fragments of MMA (kb, j), per lane (n = l / 4, c = l % 4): b0 = q[n][32 kb + 8 c + 4 j + {0,1}], b1 = .. + {2,3};
    // written by the warp that normalises query n (gv_normalize_row), zeros for absent queries
    __shared__ __align__(8) uint2 s_bq[KB * 2][32];
    __shared__ __align__(8) unsigned long long s_cmp[GV_CW][64 + GV_KEEP + 2];
    const int lane = threadIdx.x & 31, cw = threadIdx.x >> 5;
    float* s_raw = reinterpret_cast<float*>(gv_smem_raw);                                               // [GV_NQ][GV_RAW_LD]
    // the tail kernel may be scheduled as soon as every CTA has got here (it waits for this grid's completion itself)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    GV_TG(0, 0);
    GV_T(0, 1);
    // this CTA's contiguous range of 32-row tiles
    const int64_t c0 = (int64_t)q.nchunk * blockIdx.x / gridDim.x, c1 = (int64_t)q.nchunk * (blockIdx.x + 1) / gridDim.x;
    const int64_t n_mine = c1 - c0;
    // The FIRST pass is asked for with one bulk L2 prefetch per warp and tile before the query preparation (no registers, no
    // shared memory): HBM runs from the first microsecond, under the 3-4 us the queries take, and the loads of the first
    // tile loop find their lines in L2.  Every later pass is read with the loads alone (16 warps x 16 x 16 bytes per lane in
    // flight: 7.0 TB/s measured for the second pass of config 4-i) plus a two-line prefetch per row issued before each
    // selection.  Requesting the later passes up front as well was measured 2.3 us slower (tools/gv_trace.py,
    // profiles/r02_gv_trace_e*.log): the two passes are then served interleaved, the first one is complete only when 70 % of
    // everything has arrived, every load queues behind 128 MB of outstanding requests (2-3 us per group even for lines
    // that are already in L2), and all of the bank crosses the L2 slices twice.
    static_assert(GV_PASS_TILES == GV_CW, "one tile per warp and pass");
    auto prefetch_pass = [&](int64_t pass0, int len) {
        const int64_t t = pass0 + cw;
        if (lane != 0 || cw >= len || t >= n_mine) return;
        const int64_t row0 = (c0 + t) * GV_ROWS;
        const int64_t rows = q.P - row0 < GV_ROWS ? q.P - row0 : GV_ROWS;
        if (rows <= 0) return;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q.bank + row0 * q.Dp), "r"((uint32_t)(rows * q.Dp * 2)) : "memory");
    };
    // The SHORT pass comes first (n_mine mod 16 tiles, then full passes of 16): the prefetched part is small and complete
    // early (10 us on config 4-i: 10-11 of 26-27 tiles), and the passes that stream with loads alone have a tile for every
    // warp.
    const int first_len = n_mine > 0 ? (int)(n_mine - ((n_mine - 1) / GV_PASS_TILES) * GV_PASS_TILES) : 0;
    // ---- query preparation: labels, canonical normalise, B fragments ----
    auto prefetch_first = [&]() { prefetch_pass(0, first_len); };
    // every warp reads the labels itself (lane s: query s, relative to label_base, clamped): the run of its selection
    // slot needs no shared memory and no second barrier
    int32_t mylab = 0x7fffffff;
    if (lane < q.N) {
        const int32_t l = q.seg_label[lane] - q.label_base;
        mylab = l < 0 ? 0 : (l >= q.L ? q.L - 1 : l);
    }
    if (cw < GV_NQ) {
        uint2* so = &s_bq[0][0];
        if (cw < q.N) {
            const bool pub = blockIdx.x == 0;
            __nv_bfloat16* gb = pub ? q.seg_bf16 + (size_t)cw * q.Dp : nullptr;
            float* gf = (pub && q.seg_f32) ? q.seg_f32 + (size_t)cw * q.D : nullptr;
            if (q.in_dtype == SDK_IN_F16)
                gv_normalize_row<__half>(reinterpret_cast<const __half*>(q.seg_raw) + (size_t)cw * q.D, q.D, q.Dp, lane, so, cw, gb, gf, prefetch_first);
            else
                gv_normalize_row<float>(reinterpret_cast<const float*>(q.seg_raw) + (size_t)cw * q.D, q.D, q.Dp, lane, so, cw, gb, gf, prefetch_first);
        } else {
            prefetch_first();
            for (int e = lane; e < KB * 8; e += 32) so[(e >> 2) * 32 + cw * 4 + (e & 3)] = make_uint2(0u, 0u);
        }
    } else {
        prefetch_first();
    }
    if (blockIdx.x == 0 && cw == GV_NQ) {
        // label-group offsets (CSR) and label validation for the whole call: goff[g] = queries with a label < g.  Equal to
        // the usual offsets for sorted labels, monotone and inside [0, N] for anything else (which is flagged).
        int32_t f = 0, prev = -0x7fffffff;
        int32_t labs[GV_NQ];
#pragma unroll
        for (int s = 0; s < GV_NQ; ++s) labs[s] = s < q.N ? q.seg_label[s] - q.label_base : 0x7fffffff;
#pragma unroll
        for (int s = 0; s < GV_NQ; ++s) {
            if (s < q.N) {
                if (labs[s] < 0 || labs[s] >= q.L) f |= 1;
                if (labs[s] < prev) f |= 2;
                prev = labs[s];
            }
        }
        for (int g = lane; g <= q.L; g += 32) {
            int cnt = 0;
#pragma unroll
            for (int s = 0; s < GV_NQ; ++s) cnt += labs[s] < g ? 1 : 0;
            q.goff[g] = cnt;
        }
        if (f && lane == 0) atomicOr(q.flags, f);
    }
    __syncthreads();
    GV_T(0, 2);
    GV_T(0, 3);
    // ---- selection state of this warp's (query slot, half): kept list in lanes 0..15, bound on everything dropped ----
    const int sel_q = cw >> 1, sel_part = cw & 1;
    int sel_len = 0;                                               // > 0: query sel_q starts a run of that many queries of one label
    {
        const int32_t lab_q = __shfl_sync(0xffffffffu, mylab, sel_q), lab_p = __shfl_sync(0xffffffffu, mylab, sel_q > 0 ? sel_q - 1 : 0);
        const uint32_t eq = __ballot_sync(0xffffffffu, lane < q.N && lane >= sel_q && mylab == lab_q);
        if (sel_q < q.N && (sel_q == 0 || lab_p != lab_q)) sel_len = __ffs((int)~(eq >> sel_q)) - 1;
    }
    unsigned long long kept = 0ull;
    float bound = -3.0e38f;
    const int g8 = lane >> 2, c4 = lane & 3;
    const uint4* bank = reinterpret_cast<const uint4*>(q.bank);
    const int64_t pitch16 = q.Dp >> 3;                                 // row pitch in 16-byte units
#ifdef SDK_GV_TRACE
    int tpass = 0;
#endif
    // the first load group (two lines of every row; lane r: row r) of this warp's tile of the NEXT pass, issued before each
    // selection.  Plain prefetch instructions: a bulk prefetch per lane kept the warp busy for a microsecond.  Four lines, or
    // issuing them before the MMA loop instead, measured no better (profiles/r02_gv_trace_g*.log).
    auto prefetch_next = [&](int64_t pass0, int plen_) {
        const int64_t t = pass0 + plen_ + cw;
        const int64_t row = (c0 + t) * GV_ROWS + lane;
        if (t < n_mine && row < q.P) {
            const char* a = reinterpret_cast<const char*>(q.bank + row * q.Dp);
#pragma unroll
            for (int l = 0; l < 2; ++l)
                if (q.Dp * 2 > 128 * l) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128 * l));
        }
    };
    int plen = first_len;
    for (int64_t pass0 = 0; pass0 < n_mine; pass0 += plen, plen = GV_PASS_TILES) {
        const int npc = plen;                                          // (first_len + a multiple of 16 = n_mine)
        for (int j = cw; j < npc; j += GV_CW) {
            const int64_t row0 = (c0 + pass0 + j) * GV_ROWS;
            // the four bank rows of this lane (two per 16-row tile); rows past P are read from the last row and dropped later
            const uint4* rp[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                int64_t r = row0 + (h >> 1) * 16 + (h & 1) * 8 + g8;
                r = r < q.P ? r : q.P - 1;
                rp[h] = bank + r * pitch16 + c4;
            }
            float acc[2][4];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[rt][e] = 0.f;
#pragma unroll 1
            for (int kg = 0; kg < KB; kg += 4) {
                uint4 v[4][4];                                         // [k block of the group][row h]
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int h = 0; h < 4; ++h)
                        v[kb][h] = (kg + kb < KB) ? __ldcs(rp[h] + (kg + kb) * 4) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    if (kg + kb < KB) {
                        const uint2 b0 = s_bq[(kg + kb) * 2][lane], b1 = s_bq[(kg + kb) * 2 + 1][lane];
#pragma unroll
                        for (int rt = 0; rt < 2; ++rt) {
                            const uint4 lo = v[kb][rt * 2], hi = v[kb][rt * 2 + 1];
                            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(acc[rt][0]), "+f"(acc[rt][1]), "+f"(acc[rt][2]), "+f"(acc[rt][3])
                                         : "r"(lo.x), "r"(hi.x), "r"(lo.y), "r"(hi.y), "r"(b0.x), "r"(b0.y));
                            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(acc[rt][0]), "+f"(acc[rt][1]), "+f"(acc[rt][2]), "+f"(acc[rt][3])
                                         : "r"(lo.z), "r"(hi.z), "r"(lo.w), "r"(hi.w), "r"(b1.x), "r"(b1.y));
                        }
                    }
                }
            }
            // fragment (row = lane/4 (+8), queries 2*(lane%4), +1) -> s_raw[query][row of the pass]
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                const int r = j * GV_ROWS + rt * 16 + g8, cq = 2 * c4;
                s_raw[cq * GV_RAW_LD + r] = acc[rt][0];
                s_raw[(cq + 1) * GV_RAW_LD + r] = acc[rt][1];
                s_raw[cq * GV_RAW_LD + r + 8] = acc[rt][2];
                s_raw[(cq + 1) * GV_RAW_LD + r + 8] = acc[rt][3];
            }
        }
#ifdef SDK_GV_TRACE
        if (tpass < 3) GV_T(0, 4 + 4 * tpass);
#endif
        __syncthreads();                                               // the raw scores of the pass are complete
#ifdef SDK_GV_TRACE
        if (tpass < 3) GV_T(0, 5 + 4 * tpass);
#endif
        prefetch_next(pass0, plen);                                   // the DRAM works on it while the CTA selects
        if (sel_len > 0) {
            // this warp's half of the pass: tiles [sel_part * 8, +8); lane <-> rows lane + 32 t.  Pool the label's queries
            // (ascending), keep the GV_KEEP largest of (new rows, previously kept) by (score desc, row asc)
            constexpr int T = GV_PASS_TILES / GV_PARTS;
            uint32_t vk[T];
            // rows of this pass that exist: local index (tile * 32 + lane) < lim
            const int64_t left = q.P - (c0 + pass0) * GV_ROWS;
            const int lim = (int)(left < (int64_t)npc * GV_ROWS ? left : (int64_t)npc * GV_ROWS);
            const int rl0 = sel_part * T * GV_ROWS + lane;
            const float* v0 = s_raw + (size_t)sel_q * GV_RAW_LD + rl0;
            const float tau = q.tau;
            if (sel_len == 1) {                                        // one query per label (pooled centroids): no pooling loop
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const float val = v0[t * GV_ROWS];
                    vk[t] = (rl0 + t * GV_ROWS < lim && val >= tau) ? sdk_fkey(val) : 0u;
                }
            } else if (q.pool == 0) {
                const float rlen = 1.0f / (float)sel_len;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    float a = 0.f;
                    for (int s = 0; s < sel_len; ++s) a += v0[t * GV_ROWS + s * GV_RAW_LD];
                    const float val = a * rlen;
                    vk[t] = (rl0 + t * GV_ROWS < lim && val >= tau) ? sdk_fkey(val) : 0u;
                }
            } else {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    float a = -3.0e38f;
                    for (int s = 0; s < sel_len; ++s) a = fmaxf(a, v0[t * GV_ROWS + s * GV_RAW_LD]);
                    vk[t] = (rl0 + t * GV_ROWS < lim && a >= tau) ? sdk_fkey(a) : 0u;
                }
            }
            const uint32_t rbase = (uint32_t)((c0 + pass0 + sel_part * (GV_PASS_TILES / GV_PARTS)) * GV_ROWS + lane);
            GV_T(0, 20);
            kept = gv_select_top(vk, rbase, kept, bound, s_cmp[cw], lane);
        }
#ifdef SDK_GV_TRACE
        if (tpass < 3) GV_T(0, 6 + 4 * tpass);
#endif
        __syncthreads();                                               // s_raw is rewritten by the next pass
#ifdef SDK_GV_TRACE
        if (tpass < 3) GV_T(0, 7 + 4 * tpass);
        ++tpass;
#endif
    }
    // ---- publish this warp's list (every entry of every (query slot, list) is written: unused ones as 0) ----
    {
        const int64_t li = (int64_t)sel_q * q.nslots + (int64_t)blockIdx.x * GV_PARTS + sel_part;
        if (lane == 0) q.slot_bound[li] = bound;
        if (lane < GV_KEEP) q.slot_key[li * GV_KEEP + lane] = kept;
    }
    GV_T(0, 16);
    GV_TG(0, 17);
}

// ---- tail: merge + canonical re-score + select + certificate, one CTA per label ------------------------------------
struct GvTail {
    const int64_t* goff;
    int32_t N, L, k, pool, ncand, nslots, D, pitch, is_bf16;
    int32_t stage16;              // 16-byte pieces of the fp64 operand area at the head of the dynamic shared memory
    double threshold;
    float eps;
    const float* slot_bound;
    const unsigned long long* slot_key;
    const void* bank_ops;
    const void* seg_ops;
    const int32_t* row_speaker;
    const uint8_t* row_trust;
    int64_t row_offset;
    int32_t* cand_row;            // [L, ncand] diagnostics (sdk_stage_a_fetch) + the fall-back's view of stage A
    float* cand_val;
    float* gbound;
    int32_t* fb_count;
    int32_t* fb_list;
    int64_t* o_row;
    float* o_score;
    int32_t* o_count;
    uint8_t* o_trust;
    int32_t* o_spk;
};

// one fp64 fma chain over ascending d (the canonical order); zero padding beyond D adds exact zeros
template <bool BF16>
__device__ __forceinline__ double gv_dot(const void* __restrict__ a_ops, int64_t arow, const void* __restrict__ b_ops, int64_t brow,
                                         int32_t pitch, int32_t D) {
    double acc = 0.0;
    if (BF16) {
        const uint4* a = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a_ops) + arow * (int64_t)pitch);
        const uint4* b = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(b_ops) + brow * (int64_t)pitch);
        const int n8 = (D + 7) >> 3;                           // pitch is a multiple of 64 and zero padded
        uint4 pa = __ldg(a), pb = __ldg(b);
        for (int i = 0; i < n8; ++i) {
            const uint4 ca = pa, cb = pb;
            if (i + 1 < n8) { pa = __ldg(a + i + 1); pb = __ldg(b + i + 1); }
            const uint32_t wa[4] = {ca.x, ca.y, ca.z, ca.w}, wb[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                acc = fma((double)__uint_as_float(wa[h] << 16), (double)__uint_as_float(wb[h] << 16), acc);
                acc = fma((double)__uint_as_float(wa[h] & 0xffff0000u), (double)__uint_as_float(wb[h] & 0xffff0000u), acc);
            }
        }
    } else {
        const float* a = reinterpret_cast<const float*>(a_ops) + arow * (int64_t)pitch;
        const float* b = reinterpret_cast<const float*>(b_ops) + brow * (int64_t)pitch;
        int d = 0;
        if ((pitch & 3) == 0) {
            for (; d + 4 <= D; d += 4) {
                const float4 fa = __ldg(reinterpret_cast<const float4*>(a + d)), fb = __ldg(reinterpret_cast<const float4*>(b + d));
                acc = fma((double)fa.x, (double)fb.x, acc);
                acc = fma((double)fa.y, (double)fb.y, acc);
                acc = fma((double)fa.z, (double)fb.z, acc);
                acc = fma((double)fa.w, (double)fb.w, acc);
            }
        }
        for (; d < D; ++d) acc = fma((double)__ldg(a + d), (double)__ldg(b + d), acc);
    }
    return acc;
}

// ---- stage B operands in shared memory: WIDENED TO fp64 WHILE THEY ARE STAGED ------------------------------------------
// The canonical chain is one fma per element, in order, so its floor is the DFMA latency times D.  Converting inside the
// chain put two F2F.F64.F32 in front of every DFMA; that unit takes 8 cycles per warp instruction, and with all pairs of a
// small label in one warp the chain ran at 25 cycles per element (6.7 us for D = 512: tools/gv_trace.py).  The conversions
// are exact whoever does them, so the 256 threads that copy the rows do them, in parallel, and the chain is LDS + DFMA.
template <bool BF16>
__device__ __forceinline__ void gv_widen_store(double* __restrict__ row, int i, const uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (BF16) {
        double2* d = reinterpret_cast<double2*>(row + 8 * i);          // piece i = elements 8 i .. 8 i + 7
#pragma unroll
        for (int h = 0; h < 4; ++h)
            d[h] = make_double2((double)__uint_as_float(w[h] << 16), (double)__uint_as_float(w[h] & 0xffff0000u));
    } else {
        double2* d = reinterpret_cast<double2*>(row + 4 * i);          // piece i = elements 4 i .. 4 i + 3
        d[0] = make_double2((double)__uint_as_float(w[0]), (double)__uint_as_float(w[1]));
        d[1] = make_double2((double)__uint_as_float(w[2]), (double)__uint_as_float(w[3]));
    }
}
// one fp64 fma chain over ascending d (the canonical order); n2 = pairs of elements (an odd D ends in a zero pair half)
__device__ __forceinline__ double gv_dot_f64(const double* __restrict__ a, const double* __restrict__ b, int n2) {
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    double acc = 0.0;
#pragma unroll 8
    for (int i = 0; i < n2; ++i) {
        const double2 x = a2[i], y = b2[i];
        acc = fma(x.x, y.x, acc);
        acc = fma(x.y, y.y, acc);
    }
    return acc;
}

#define GV_TAIL_CHUNK 16             // candidate rows staged at a time
#define GV_TAIL_LB 10                // 16-byte list pieces per thread in flight (2 x 148 lists x 16 keys = 2368 pieces: one batch)
#define GV_TAIL_RB 4                 // 16-byte operand pieces per thread in flight (16 rows x 1 KB = 1024 pieces: one batch)
__global__ void __launch_bounds__(GV_TAIL_THREADS)
k_gemv8_tail(const __grid_constant__ GvTail p) {
    extern __shared__ uint4 s_stage[];                         // fp64 operands [GV_NQ segment rows + GV_TAIL_CHUNK bank rows][pitch + 2], then the lists
    __shared__ unsigned long long s_key[GV_TAIL_CAP];          // compacted candidates
    __shared__ unsigned long long s_sel[64];                   // the ncand best, by rank
    __shared__ long long s_pool[64];
    __shared__ int s_rank[128];
    __shared__ unsigned long long s_T0, s_Tsel;
    __shared__ int s_m, s_over;
    __shared__ float s_bnd[GV_TAIL_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;
    const int k = p.k, ncand = p.ncand, nslots = p.nslots;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // sdk_assign's kernel may queue up behind this one
    GV_TG(1, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");                // everything the stream kernel wrote is visible
    GV_TG(1, 1);
    GV_T(1, 2);
    int64_t s0 = p.goff[g], s1 = p.goff[g + 1];
    s0 = s0 < 0 ? 0 : (s0 > p.N ? p.N : s0);
    s1 = s1 < 0 ? 0 : (s1 > p.N ? p.N : s1);
    const int n = (int)(s1 - s0);
    for (int i = tid; i < k; i += GV_TAIL_THREADS) {
        p.o_row[(int64_t)g * k + i] = -1;
        p.o_score[(int64_t)g * k + i] = 0.f;
        p.o_trust[(int64_t)g * k + i] = SDK_TRUST_UNKNOWN;
        p.o_spk[(int64_t)g * k + i] = -1;
    }
    for (int i = tid; i < ncand; i += GV_TAIL_THREADS) {
        p.cand_row[(int64_t)g * ncand + i] = -1;
        p.cand_val[(int64_t)g * ncand + i] = 0.f;
    }
    if (n <= 0) {
        if (tid == 0) { p.o_count[g] = 0; p.gbound[g] = -3.0e38f; }
        return;
    }
    const int64_t lbase = (int64_t)s0 * nslots;                // lists of the label: query slot = its first query
    // operand geometry: rows of `pitch` elements (bf16: Dp, zero padded; fp32: D), copied in 16-byte pieces
    const int row_bytes = p.is_bf16 ? p.pitch * 2 : p.pitch * 4;
    const bool staged = (row_bytes & 15) == 0;
    const int row16 = row_bytes >> 4;
    const int ldd = p.pitch + 2;                               // + 16 bytes: rows of different candidates in different banks
    double* s_seg = reinterpret_cast<double*>(s_stage);
    double* s_row = s_seg + (size_t)GV_NQ * ldd;
    unsigned long long* s_lkey = reinterpret_cast<unsigned long long*>(s_stage + p.stage16);   // [nslots][GV_KEEP] keys, 0 = no entry
    float* s_lbound = reinterpret_cast<float*>(s_lkey + (size_t)nslots * GV_KEEP);            // [nslots]
    if (tid == 0) { s_m = 0; s_over = 0; s_T0 = 0ull; s_Tsel = 0ull; }
    if (tid < 128) s_rank[tid] = 0;
    // ---- 0. everything this CTA needs from the stream kernel is REQUESTED before anything is used: the label's segment
    //         operands, its lists (one coalesced sweep of 16-byte pieces) and their bounds; then stored (the lists as they
    //         are, the operands widened to fp64).  Read list by list this phase cost a memory round trip per entry. ----
    const uint4* gseg = reinterpret_cast<const uint4*>(p.seg_ops) + (int64_t)s0 * row16;      // the label's rows are contiguous
    const int nseg16 = staged ? n * row16 : 0;
    uint4 sb[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int idx = tid + u * GV_TAIL_THREADS;
        if (idx < nseg16) sb[u] = __ldg(gseg + idx);
    }
    float bb[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int sidx = tid + u * GV_TAIL_THREADS;
        bb[u] = sidx < nslots ? p.slot_bound[lbase + sidx] : -3.0e38f;
    }
    {
        const ulonglong2* gk = reinterpret_cast<const ulonglong2*>(p.slot_key + lbase * GV_KEEP);
        ulonglong2* sk2 = reinterpret_cast<ulonglong2*>(s_lkey);
        const int nl16 = nslots * (GV_KEEP / 2);
        for (int base = 0; base < nl16; base += GV_TAIL_THREADS * GV_TAIL_LB) {
            ulonglong2 lb[GV_TAIL_LB];
#pragma unroll
            for (int u = 0; u < GV_TAIL_LB; ++u) {
                const int idx = base + u * GV_TAIL_THREADS + tid;
                if (idx < nl16) lb[u] = gk[idx];
            }
#pragma unroll
            for (int u = 0; u < GV_TAIL_LB; ++u) {
                const int idx = base + u * GV_TAIL_THREADS + tid;
                if (idx < nl16) sk2[idx] = lb[u];
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int sidx = tid + u * GV_TAIL_THREADS;
        if (sidx < nslots) s_lbound[sidx] = bb[u];
    }
    for (int sidx = tid + 2 * GV_TAIL_THREADS; sidx < nslots; sidx += GV_TAIL_THREADS) s_lbound[sidx] = p.slot_bound[lbase + sidx];
    for (int base = 0; base < nseg16; base += 2 * GV_TAIL_THREADS) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int idx = base + tid + u * GV_TAIL_THREADS;
            if (idx < nseg16) {
                if (base > 0) sb[u] = __ldg(gseg + idx);
                const int t = idx / row16, i = idx - t * row16;
                if (p.is_bf16) gv_widen_store<true>(s_seg + (size_t)t * ldd, i, sb[u]);
                else gv_widen_store<false>(s_seg + (size_t)t * ldd, i, sb[u]);
            }
        }
    }
    __syncthreads();
    GV_T(1, 3);
    // ---- 1. T0 = ncand-th largest list maximum: a lower bound on the ncand-th best entry overall ----
    // (any lower bound will do: the maxima of the first 128 lists are ranked, not all 2 x grid of them; two threads per list)
    const int nmax = nslots < 128 ? nslots : 128;
    {
        const int sidx = tid & 127, half = tid >> 7, mid = (nmax + 1) >> 1;
        if (sidx < nmax) {
            const unsigned long long mine = s_lkey[sidx * GV_KEEP];
            if (mine != 0ull) {
                const int j0 = half ? mid : 0, j1 = half ? nmax : mid;
                int rank = 0;
#pragma unroll 4
                for (int j = j0; j < j1; ++j) rank += s_lkey[j * GV_KEEP] > mine ? 1 : 0;
                if (rank) atomicAdd(&s_rank[sidx], rank);
            }
        }
    }
    __syncthreads();
    if (tid < nmax) {
        const unsigned long long mine = s_lkey[tid * GV_KEEP];
        if (mine != 0ull && s_rank[tid] == ncand - 1) s_T0 = mine;  // keys are unique: exactly one list has this rank (if any)
    }
    __syncthreads();
    GV_T(1, 4);
    // ---- 2. compaction of every entry >= T0 (each list is sorted: a prefix) ----
    const unsigned long long T0 = s_T0;
    for (int sidx = tid; sidx < nslots; sidx += GV_TAIL_THREADS) {
        for (int e = 0; e < GV_KEEP; ++e) {
            const unsigned long long key = s_lkey[sidx * GV_KEEP + e];
            if (key == 0ull || key < T0) break;
            const int pos = atomicAdd(&s_m, 1);
            if (pos < GV_TAIL_CAP) s_key[pos] = key; else s_over = 1;
        }
    }
    __syncthreads();
    GV_T(1, 5);
    const int m = s_m < GV_TAIL_CAP ? s_m : GV_TAIL_CAP;
    // ---- 3. exact rank of the compacted entries; the ncand best in order ----
    for (int i = tid; i < 64; i += GV_TAIL_THREADS) s_sel[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < m; i += GV_TAIL_THREADS) {
        const unsigned long long mine = s_key[i];
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < m; ++j) rank += s_key[j] > mine ? 1 : 0;
        if (rank < ncand) s_sel[rank] = mine;
        if (rank == ncand - 1) s_Tsel = mine;
    }
    __syncthreads();
    GV_T(1, 6);
    const int nc = m < ncand ? m : ncand;
    const unsigned long long Tsel = s_Tsel;                        // 0: every entry of every list was selected
    // the first chunk of candidate rows is requested NOW and stored after the bound phase; every candidate's speaker /
    // trust too (used after stage B: the extraction loop at the end then never waits for memory)
    const uint4* gbank = reinterpret_cast<const uint4*>(p.bank_ops);
    uint4 rb[GV_TAIL_RB];
    {
        const int cc0 = nc < GV_TAIL_CHUNK ? nc : GV_TAIL_CHUNK;
        const int np0 = staged ? cc0 * row16 : 0;
#pragma unroll
        for (int u = 0; u < GV_TAIL_RB; ++u) {
            const int idx = tid + u * GV_TAIL_THREADS;
            if (idx < np0) {
                const int j = idx / row16, i = idx - j * row16;
                const int64_t row = (int64_t)(0xffffffffu - (uint32_t)s_sel[j]);
                rb[u] = __ldg(gbank + row * row16 + i);
            }
        }
    }
    int32_t spk0 = -1, spk1 = -1;
    uint32_t tr0 = SDK_TRUST_UNKNOWN, tr1 = SDK_TRUST_UNKNOWN;
    if (warp == 0) {
        if (lane < nc) {
            const uint32_t row = 0xffffffffu - (uint32_t)s_sel[lane];
            spk0 = p.row_speaker[row];
            if (p.row_trust) tr0 = p.row_trust[row];
        }
        if (lane + 32 < nc) {
            const uint32_t row = 0xffffffffu - (uint32_t)s_sel[lane + 32];
            spk1 = p.row_speaker[row];
            if (p.row_trust) tr1 = p.row_trust[row];
        }
    }
    // ---- 4. bound on everything that is not a candidate: dropped inside the stream kernel, or left in a list ----
    float bnd = -3.0e38f;
    for (int sidx = tid; sidx < nslots; sidx += GV_TAIL_THREADS) {
        bnd = fmaxf(bnd, s_lbound[sidx]);
        if (Tsel != 0ull) {
            for (int e = 0; e < GV_KEEP; ++e) {
                const unsigned long long key = s_lkey[sidx * GV_KEEP + e];
                if (key == 0ull) break;
                if (key < Tsel) { bnd = fmaxf(bnd, sdk_funkey((uint32_t)(key >> 32))); break; }
            }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) bnd = fmaxf(bnd, __shfl_xor_sync(0xffffffffu, bnd, off));
    if (lane == 0) s_bnd[warp] = bnd;
    if (tid < 64) s_pool[tid] = p.pool == 0 ? 0ll : LLONG_MIN;
    __syncthreads();
    GV_T(1, 7);
    if (tid == 0) {
        for (int w = 1; w < GV_TAIL_THREADS / 32; ++w) bnd = fmaxf(bnd, s_bnd[w]);
        if (s_over) bnd = 3.0e38f;                                 // more live entries than the tail holds: cannot certify
        p.gbound[g] = bnd;
        s_bnd[0] = bnd;
    }
    for (int i = tid; i < nc; i += GV_TAIL_THREADS) {
        p.cand_row[(int64_t)g * ncand + i] = (int32_t)(0xffffffffu - (uint32_t)s_sel[i]);
        p.cand_val[(int64_t)g * ncand + i] = sdk_funkey((uint32_t)(s_sel[i] >> 32));
    }
    // ---- 5. stage B: canonical pooled scores of the candidates (thread = (candidate, segment) pair) over the fp64 rows ----
    if (staged) {
        const int n2 = p.is_bf16 ? (p.D + 1) >> 1 : p.D >> 1;      // (bf16 rows are zero padded past D; fp32 rows here have D % 4 == 0)
        for (int cb = 0; cb < nc; cb += GV_TAIL_CHUNK) {
            const int cc = nc - cb < GV_TAIL_CHUNK ? nc - cb : GV_TAIL_CHUNK;
            if (cb > 0) __syncthreads();                           // the previous chunk has been consumed
            for (int base = 0; base < cc * row16; base += GV_TAIL_RB * GV_TAIL_THREADS) {
#pragma unroll
                for (int u = 0; u < GV_TAIL_RB; ++u) {
                    const int idx = base + tid + u * GV_TAIL_THREADS;
                    if (idx < cc * row16 && (cb > 0 || base > 0)) {
                        const int j = idx / row16, i = idx - j * row16;
                        const int64_t row = (int64_t)(0xffffffffu - (uint32_t)s_sel[cb + j]);
                        rb[u] = __ldg(gbank + row * row16 + i);
                    }
                }
#pragma unroll
                for (int u = 0; u < GV_TAIL_RB; ++u) {
                    const int idx = base + tid + u * GV_TAIL_THREADS;
                    if (idx < cc * row16) {
                        const int j = idx / row16, i = idx - j * row16;
                        if (p.is_bf16) gv_widen_store<true>(s_row + (size_t)j * ldd, i, rb[u]);
                        else gv_widen_store<false>(s_row + (size_t)j * ldd, i, rb[u]);
                    }
                }
            }
            __syncthreads();
            if (cb == 0) GV_T(1, 8);
            for (int pr = tid; pr < cc * n; pr += GV_TAIL_THREADS) {
                const int j = pr / n, t = pr - j * n;
                const double sc = gv_dot_f64(s_seg + (size_t)t * ldd, s_row + (size_t)j * ldd, n2);
                const long long qv = __double2ll_rn(sc * SDK_Q30);
                if (p.pool == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&s_pool[cb + j]), (unsigned long long)qv);
                else atomicMax(&s_pool[cb + j], qv);
            }
        }
    } else {
        for (int pr = tid; pr < nc * n; pr += GV_TAIL_THREADS) {   // fp32 rows that are not 16-byte multiples: straight from global
            const int j = pr / n, t = pr - j * n;
            const int64_t row = (int64_t)(0xffffffffu - (uint32_t)s_sel[j]);
            const double sc = gv_dot<false>(p.seg_ops, s0 + t, p.bank_ops, row, p.pitch, p.D);
            const long long qv = __double2ll_rn(sc * SDK_Q30);
            if (p.pool == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&s_pool[j]), (unsigned long long)qv);
            else atomicMax(&s_pool[j], qv);
        }
    }
    __syncthreads();
    GV_T(1, 9);
    // ---- 6. row -> speaker max, threshold, ordered top-k, certificate (k_select's semantics), one warp, no serial
    //         extraction: a candidate is a match iff it passes the threshold and no passing candidate of the same speaker has
    //         a larger (score, row) key; its position is the number of matches with a larger key (the first k are written).
    //         Same result as taking the candidates in key order, skipping speakers already taken, stopping at the first
    //         score below the threshold or at k matches -- scores do not increase along that order. ----
    if (warp != 0) return;
    unsigned long long kp0 = 0ull, kp1 = 0ull;                     // keys of the lane's candidates that pass the threshold
    if (lane < nc) {
        const float sim = sdk_pool_finish(s_pool[lane], n, p.pool);
        if ((double)sim >= p.threshold) kp0 = ((unsigned long long)sdk_fkey(sim) << 32) | (uint32_t)s_sel[lane];
    }
    if (lane + 32 < nc) {
        const float sim = sdk_pool_finish(s_pool[lane + 32], n, p.pool);
        if ((double)sim >= p.threshold) kp1 = ((unsigned long long)sdk_fkey(sim) << 32) | (uint32_t)s_sel[lane + 32];
    }
    const int n_lo = nc < 32 ? nc : 32;
    bool d0 = false, d1 = false;
    for (int j = 0; j < n_lo; ++j) {
        const unsigned long long kj = __shfl_sync(0xffffffffu, kp0, j);
        const int32_t sj = __shfl_sync(0xffffffffu, spk0, j);
        d0 |= kj > kp0 && sj == spk0;
        d1 |= kj > kp1 && sj == spk1;
    }
    for (int j = 32; j < nc; ++j) {
        const unsigned long long kj = __shfl_sync(0xffffffffu, kp1, j - 32);
        const int32_t sj = __shfl_sync(0xffffffffu, spk1, j - 32);
        d0 |= kj > kp0 && sj == spk0;
        d1 |= kj > kp1 && sj == spk1;
    }
    const unsigned long long ku0 = d0 ? 0ull : kp0, ku1 = d1 ? 0ull : kp1;                  // the matches
    int r0 = 0, r1 = 0;
    for (int j = 0; j < n_lo; ++j) {
        const unsigned long long kj = __shfl_sync(0xffffffffu, ku0, j);
        r0 += kj > ku0 ? 1 : 0;
        r1 += kj > ku1 ? 1 : 0;
    }
    for (int j = 32; j < nc; ++j) {
        const unsigned long long kj = __shfl_sync(0xffffffffu, ku1, j - 32);
        r0 += kj > ku0 ? 1 : 0;
        r1 += kj > ku1 ? 1 : 0;
    }
    const int nu = __popc(__ballot_sync(0xffffffffu, ku0 != 0ull)) + __popc(__ballot_sync(0xffffffffu, ku1 != 0ull));
    const int cnt = nu < k ? nu : k;
    if (ku0 != 0ull && r0 < k) {
        p.o_row[(int64_t)g * k + r0] = (int64_t)(int32_t)(0xffffffffu - (uint32_t)ku0) + p.row_offset;
        p.o_score[(int64_t)g * k + r0] = sdk_funkey((uint32_t)(ku0 >> 32));
        p.o_trust[(int64_t)g * k + r0] = (uint8_t)tr0;
        p.o_spk[(int64_t)g * k + r0] = spk0;
    }
    if (ku1 != 0ull && r1 < k) {
        p.o_row[(int64_t)g * k + r1] = (int64_t)(int32_t)(0xffffffffu - (uint32_t)ku1) + p.row_offset;
        p.o_score[(int64_t)g * k + r1] = sdk_funkey((uint32_t)(ku1 >> 32));
        p.o_trust[(int64_t)g * k + r1] = (uint8_t)tr1;
        p.o_spk[(int64_t)g * k + r1] = spk1;
    }
    // the k-th match's score (only read when there are k of them)
    float kth = 0.f;
    {
        const uint32_t m0 = __ballot_sync(0xffffffffu, ku0 != 0ull && r0 == k - 1), m1 = __ballot_sync(0xffffffffu, ku1 != 0ull && r1 == k - 1);
        const float f0 = sdk_funkey((uint32_t)(ku0 >> 32)), f1 = sdk_funkey((uint32_t)(ku1 >> 32));
        if (m0) kth = __shfl_sync(0xffffffffu, f0, __ffs(m0) - 1);
        else if (m1) kth = __shfl_sync(0xffffffffu, f1, __ffs(m1) - 1);
    }
    if (lane == 0) {
        p.o_count[g] = cnt;
        // certificate (select.cu): every row that was not re-scored has a stage-A score <= bound, hence a canonical score
        // <= bound + eps; it cannot enter the result if that is below the threshold, or below the k-th kept score
        const double b = (double)s_bnd[0] + (double)p.eps;
        const bool safe = (b < p.threshold) || (cnt == k && b < (double)kth);
        if (!safe) {
            const int pos = atomicAdd(p.fb_count, 1);
            p.fb_list[pos] = g;
        }
    }
    GV_T(1, 10);
    GV_TG(1, 11);
}

int sdk_gemv_applicable(int64_t N, int32_t Dp) { return N >= 1 && N <= GV_NQ && sdk_poolgemm_supported(Dp); }

// The whole small-query identify: stream kernel + tail.  Everything it needs is reserved here; nothing is synchronised.
int sdk_launch_gemv_identify(sdk_ctx* c, const void* d_seg_raw, int32_t in_dtype, const int32_t* d_seg_label, int32_t label_base, int64_t N,
                             int32_t L, int32_t pool, double threshold, int32_t k, float tau, float eps, int32_t ncand, int32_t* d_flags,
                             int64_t* o_row, float* o_score, int32_t* o_count, uint8_t* o_trust, int32_t* o_spk) {
    const int32_t D = c->D, Dp = c->Dp;
    const int64_t P = c->P;
    const bool bf16 = c->dtype == SDK_DTYPE_BF16;
    if (!sdk_gemv_applicable(N, Dp)) return sdk_fail(c, SDK_EINVAL, "gemv path: needs 1..8 segments and a supported D");
    if (P > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "gemv path: at most 2^31-1 bank rows");
    if (ncand > 64) ncand = 64;
    const int32_t nchunk = (int32_t)((P + GV_ROWS - 1) / GV_ROWS);
    const int grid = (int)std::min<int64_t>(nchunk, std::min(c->sm_count, 148));
    const int32_t nslots = grid * GV_PARTS;
    const size_t smem = (size_t)GV_NQ * GV_RAW_LD * 4;
    SDK_TRY(sdk_reserve(c, c->goff, (size_t)(L + 1) * 8));
    SDK_TRY(sdk_reserve(c, c->seg_bf16, (size_t)N * Dp * 2));
    if (!bf16) SDK_TRY(sdk_reserve(c, c->seg_f32, (size_t)N * D * 4));
    SDK_TRY(sdk_reserve(c, c->slot_bound, (size_t)GV_NQ * nslots * 4));
    SDK_TRY(sdk_reserve(c, c->slot_row, (size_t)GV_NQ * nslots * GV_KEEP * 8));          // 64-bit keys on this path
    SDK_TRY(sdk_reserve(c, c->cand_row, (size_t)L * ncand * 4));
    SDK_TRY(sdk_reserve(c, c->cand_val, (size_t)L * ncand * 4));
    SDK_TRY(sdk_reserve(c, c->gbound, (size_t)L * 4));
    SDK_TRY(sdk_reserve(c, c->fb_list, (size_t)L * 4));
    GvParams q;
    q.seg_raw = d_seg_raw;
    q.seg_label = d_seg_label;
    q.in_dtype = in_dtype;
    q.label_base = label_base;
    q.N = (int32_t)N;
    q.L = L;
    q.D = D;
    q.Dp = Dp;
    q.P = P;
    q.pool = pool;
    q.tau = tau;
    q.nchunk = nchunk;
    q.bank = (const __nv_bfloat16*)c->bank_bf16.p;
    q.nslots = nslots;
    q.goff = (int64_t*)c->goff.p;
    q.flags = d_flags + SDK_FLAG_LABEL;
    q.seg_bf16 = (__nv_bfloat16*)c->seg_bf16.p;
    q.seg_f32 = bf16 ? nullptr : (float*)c->seg_f32.p;
    q.slot_bound = (float*)c->slot_bound.p;
    q.slot_key = (unsigned long long*)c->slot_row.p;
    {
        sdk_prof_scope ps(c, "poolgemm");         // stage A of the certified top-k, whichever kernel runs it
#define GV_LAUNCH(K)                                                                                            \
    do {                                                                                                            \
        SDK_CUDA(c, cudaFuncSetAttribute(k_gemv8<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        k_gemv8<K><<<grid, GV_THREADS, smem, c->stream>>>(q);                                                      \
    } while (0)
        switch (Dp / 32) {
            case 2: GV_LAUNCH(2); break;
            case 4: GV_LAUNCH(4); break;
            case 6: GV_LAUNCH(6); break;
            case 8: GV_LAUNCH(8); break;
            case 10: GV_LAUNCH(10); break;
            case 12: GV_LAUNCH(12); break;
            case 14: GV_LAUNCH(14); break;
            default: GV_LAUNCH(16); break;
        }
#undef GV_LAUNCH
        c->launches++;
        SDK_CUDA(c, cudaGetLastError());
    }
    GvTail t;
    t.goff = (const int64_t*)c->goff.p;
    t.N = (int32_t)N;
    t.L = L;
    t.k = k;
    t.pool = pool;
    t.ncand = ncand;
    t.nslots = nslots;
    t.D = D;
    t.pitch = bf16 ? Dp : D;
    t.is_bf16 = bf16 ? 1 : 0;
    t.threshold = threshold;
    t.eps = eps;
    t.slot_bound = q.slot_bound;
    t.slot_key = q.slot_key;
    t.bank_ops = bf16 ? c->bank_bf16.p : c->bank_f32.p;
    t.seg_ops = bf16 ? (const void*)c->seg_bf16.p : (const void*)c->seg_f32.p;
    t.row_speaker = (const int32_t*)c->row_speaker.p;
    t.row_trust = (const uint8_t*)c->row_trust.p;
    t.row_offset = c->row_offset;
    t.cand_row = (int32_t*)c->cand_row.p;
    t.cand_val = (float*)c->cand_val.p;
    t.gbound = (float*)c->gbound.p;
    t.fb_count = d_flags + SDK_FLAG_FB;
    t.fb_list = (int32_t*)c->fb_list.p;
    t.o_row = o_row;
    t.o_score = o_score;
    t.o_count = o_count;
    t.o_trust = o_trust;
    t.o_spk = o_spk;
    {
        sdk_prof_scope ps(c, "select");           // merge + canonical re-score + select + certificate in one launch
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)L);
        cfg.blockDim = dim3(GV_TAIL_THREADS);
        const int row_bytes = bf16 ? Dp * 2 : D * 4;
        const size_t stage_bytes = (row_bytes & 15) == 0 ? (size_t)(GV_NQ + GV_TAIL_CHUNK) * (size_t)(t.pitch + 2) * 8 : 16;
        t.stage16 = (int32_t)(stage_bytes / 16);
        const size_t tail_smem = stage_bytes + (size_t)nslots * (GV_KEEP * 8 + 8);
        SDK_CUDA(c, cudaFuncSetAttribute(k_gemv8_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem));
        cfg.dynamicSmemBytes = tail_smem;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        SDK_CUDA(c, cudaLaunchKernelEx(&cfg, k_gemv8_tail, t));
        c->launches++;
    }
    c->slot_g0 = c->slot_g1 = 0;                  // no candidate slots of the tcgen05 kind: a failed certificate goes to the exhaustive pass
    c->slot_nsub = 0;
    c->slot_by_col = false;
    return SDK_OK;
}

// ---- read probe: what ONE plain pass over the bank operands costs on this GPU (bench.py: the practical ceiling beside
// the nominal copy bandwidth, which was measured on a read+write copy of 2 GB, not on a 40 us read of 128 MB) ----
__global__ void __launch_bounds__(512) k_probe_read(const uint4* __restrict__ p, int64_t n16, uint32_t* __restrict__ out) {
    uint32_t acc = 0u;
    const int64_t stride = (int64_t)gridDim.x * 512;
    int64_t i = (int64_t)blockIdx.x * 512 + threadIdx.x;
    for (; i + 7 * stride < n16; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += stride) { const uint4 v = __ldcs(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x9e3779b9u) out[0] = acc;                  // (keeps the loads alive; practically never taken)
}
int sdk_launch_probe_read(sdk_ctx* c) {
    if (!c->bank_ok || c->P <= 0) return sdk_fail(c, SDK_ESTATE, "read probe: no bank loaded");
    SDK_TRY(sdk_reserve(c, c->flags, SDK_NFLAGS * 4));
    const int64_t n16 = (int64_t)c->P * c->Dp / 8;
    k_probe_read<<<c->sm_count * 2, 512, 0, c->stream>>>((const uint4*)c->bank_bf16.p, n16, (uint32_t*)c->flags.p);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}
