// normalize.cu -- K1: fused canonical L2-normalise-and-cast, plus label-group offsets.
//
// K1 is HBM-bound: one warp per row, 128-bit coalesced loads (lane j owns float4 chunks j, j+32, ..),
// the row is held in registers between the norm and the scale, and the outputs are a zero-padded
// bf16 copy (8-byte packed stores, operand of the tcgen05 GEMM and of bf16 canonical scoring) and/or
// the fp32 normalised copy (operand of fp32 canonical scoring).
// Algorithmic bytes per row: 4*D read + 2*Dp (bf16) [+ 4*D (fp32 copy)] written.
//
// The squared norm uses the canonical order of oracle/canonical.c step (1): fp64 partial sums per
// lane over ascending elements, butterfly xor 16,8,4,2,1 -- identical on every lane.
#include "common.cuh"

template <int NQ, typename TIn>
__global__ void __launch_bounds__(256) k_normalize_vec(const TIn* __restrict__ x, int64_t n, int32_t D,
                                                       int32_t Dp, float* __restrict__ f32out,
                                                       __nv_bfloat16* __restrict__ bf16out) {
    // R consecutive rows per warp iteration: all of their loads are in flight before the first norm is reduced
    // (one 768-byte row per warp does not cover the HBM latency-bandwidth product)
    constexpr int R = (NQ <= 2) ? 4 : (NQ <= 4 ? 2 : 1);
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nq = D >> 2, nqp = Dp >> 2;
    const int64_t n_runs = (n + R - 1) / R;
    for (int64_t run = warp0; run < n_runs; run += nwarps) {
        const int64_t row0 = run * R;
        float4 v[R][NQ];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool live = row0 + r < n;
            const TIn* xr = x + (live ? row0 + r : row0) * (int64_t)D;
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                int q = lane + 32 * i;
                v[r][i] = (live && q < nq) ? sdk_in<TIn>::ld4(xr, q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            if (row >= n) break;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                double a = (double)v[r][i].x, b = (double)v[r][i].y, c = (double)v[r][i].z, d = (double)v[r][i].w;
                s = fma(a, a, s);
                s = fma(b, b, s);
                s = fma(c, c, s);
                s = fma(d, d, s);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
            float nrm = (float)sqrt(s);
            float den = nrm > 1e-12f ? nrm : 1e-12f;
            const float inv = __fdiv_rn(1.0f, den);
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                int q = lane + 32 * i;
                float4 o;
                o.x = __fmul_rn(v[r][i].x, inv);
                o.y = __fmul_rn(v[r][i].y, inv);
                o.z = __fmul_rn(v[r][i].z, inv);
                o.w = __fmul_rn(v[r][i].w, inv);
                if (f32out && q < nq) reinterpret_cast<float4*>(f32out + row * (int64_t)D)[q] = o;
                if (bf16out && q < nqp) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y);   // zero beyond D (v was zero)
                    __nv_bfloat162 hi = __floats2bfloat162_rn(o.z, o.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    reinterpret_cast<uint2*>(bf16out + row * (int64_t)Dp)[q] = pk;
                }
            }
            if (bf16out) {   // padding chunks beyond what the NQ registers cover
                for (int q = lane + 32 * NQ; q < nqp; q += 32)
                    reinterpret_cast<uint2*>(bf16out + row * (int64_t)Dp)[q] = make_uint2(0u, 0u);
            }
        }
    }
}

// any D (scalar loads, two passes over the row; second pass hits L1/L2)
template <typename TIn>
__global__ void __launch_bounds__(256) k_normalize_generic(const TIn* __restrict__ x, int64_t n, int32_t D,
                                                           int32_t Dp, float* __restrict__ f32out,
                                                           __nv_bfloat16* __restrict__ bf16out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n; row += nwarps) {
        const TIn* xr = x + row * (int64_t)D;
        double s = 0.0;
        for (int q = lane; 4 * q < D; q += 32)
            for (int t = 0; t < 4; ++t) {
                int e = 4 * q + t;
                if (e < D) { double a = (double)sdk_in<TIn>::ld1(xr, e); s = fma(a, a, s); }
            }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
        float nrm = (float)sqrt(s);
        float den = nrm > 1e-12f ? nrm : 1e-12f;
        const float inv = __fdiv_rn(1.0f, den);
        for (int e = lane; e < Dp; e += 32) {
            float o = e < D ? __fmul_rn(sdk_in<TIn>::ld1(xr, e), inv) : 0.f;
            if (f32out && e < D) f32out[row * (int64_t)D + e] = o;
            if (bf16out) bf16out[row * (int64_t)Dp + e] = __float2bfloat16_rn(o);
        }
    }
}

// ---- pool-first stage A (a DIFFERENT algorithm for mean pooling, option "poolfirst"): K1 also accumulates the label
// centroids.  Mean pooling is linear: mean_t <x_t, b> = <mean_t x_t, b>, so a stage A that only has to be eps-accurate
// (stage B re-scores the candidates in the canonical arithmetic) can contract G centroids instead of N segments against
// the bank -- N/G times fewer flops, and the step becomes HBM-bound on this kernel: read 4*D (fp16 input: 2*D), write
// 2*Dp bytes per segment.  A warp owns 64 consecutive rows (R at a time in flight, as k_normalize_vec), keeps the running
// sum of the operand values (bf16-rounded for bf16 banks: the values stage B multiplies) of the current label in
// registers and flushes it with 128-bit vector atomics when the label changes: ~1.3 flushes per 64 rows on config 3.
// fp32 sums: <= 64 register adds + n/64 atomic adds per label (the certificate's margin model counts them).
#define SDK_PF_CHUNK 64
template <int NQ, typename TIn>
__global__ void __launch_bounds__(256) k_normalize_centroid(const TIn* __restrict__ x, const int32_t* __restrict__ lab, int32_t label_base,
                                                            int64_t n, int32_t D, int32_t Dp, float* __restrict__ f32out,
                                                            __nv_bfloat16* __restrict__ bf16out, float* __restrict__ csum, int32_t round_bf16) {
    constexpr int R = (NQ <= 2) ? 4 : (NQ <= 4 ? 2 : 1);
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nq = D >> 2, nqp = Dp >> 2;
    const int64_t n_chunks = (n + SDK_PF_CHUNK - 1) / SDK_PF_CHUNK;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += nwarps) {
        const int64_t row_lo = chunk * SDK_PF_CHUNK, row_hi = (row_lo + SDK_PF_CHUNK < n) ? row_lo + SDK_PF_CHUNK : n;
        float4 acc[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        int32_t cur = -1;
        auto flush = [&](int32_t g) {
            float4* dst = reinterpret_cast<float4*>(csum + (int64_t)g * Dp);
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const int q = lane + 32 * i;
                if (q < nq) atomicAdd(dst + q, acc[i]);
                acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        for (int64_t row0 = row_lo; row0 < row_hi; row0 += R) {
            float4 v[R][NQ];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool live = row0 + r < row_hi;
                const TIn* xr = x + (live ? row0 + r : row0) * (int64_t)D;
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    int q = lane + 32 * i;
                    v[r][i] = (live && q < nq) ? sdk_in<TIn>::ld4(xr, q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            int32_t my_lab = -1;
            if (lane < R && row0 + lane < row_hi) my_lab = lab[row0 + lane] - label_base;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t row = row0 + r;
                const int32_t g = __shfl_sync(0xffffffffu, my_lab, r);
                if (row >= row_hi) break;
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    double a = (double)v[r][i].x, b = (double)v[r][i].y, c = (double)v[r][i].z, d = (double)v[r][i].w;
                    s = fma(a, a, s);
                    s = fma(b, b, s);
                    s = fma(c, c, s);
                    s = fma(d, d, s);
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
                float nrm = (float)sqrt(s);
                float den = nrm > 1e-12f ? nrm : 1e-12f;
                const float inv = __fdiv_rn(1.0f, den);
                if (g != cur) {
                    if (cur >= 0) flush(cur);
                    cur = g;
                }
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    int q = lane + 32 * i;
                    float4 o;
                    o.x = __fmul_rn(v[r][i].x, inv);
                    o.y = __fmul_rn(v[r][i].y, inv);
                    o.z = __fmul_rn(v[r][i].z, inv);
                    o.w = __fmul_rn(v[r][i].w, inv);
                    if (f32out && q < nq) reinterpret_cast<float4*>(f32out + row * (int64_t)D)[q] = o;
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y);   // zero beyond D (v was zero)
                    __nv_bfloat162 hi = __floats2bfloat162_rn(o.z, o.w);
                    if (q < nqp) {
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&lo);
                        pk.y = *reinterpret_cast<uint32_t*>(&hi);
                        reinterpret_cast<uint2*>(bf16out + row * (int64_t)Dp)[q] = pk;
                    }
                    if (round_bf16) {
                        const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
                        o = make_float4(a.x, a.y, b.x, b.y);
                    }
                    acc[i].x += o.x; acc[i].y += o.y; acc[i].z += o.z; acc[i].w += o.w;
                }
                for (int q = lane + 32 * NQ; q < nqp; q += 32) reinterpret_cast<uint2*>(bf16out + row * (int64_t)Dp)[q] = make_uint2(0u, 0u);
            }
        }
        if (cur >= 0) flush(cur);
    }
}

// centroid = sum / n, split into two bf16 rows (hi + lo: 16 significant bits), both doubled so that the MEAN over the two
// "segments" of a label gives <hi, b> + <lo, b>.  The rows are written straight into the accumulate-pooling layout
// (poolacc.cu) together with its plan: one accumulator column per label, blocks of 256 labels, two steps per block --
// row of (label g, half t) = ((g / 256) * 2 + t) * 256 + g % 256; the unused columns of the last block are zero rows.
// goff2[g] = 2 g (every label "has" two segments).
#define SDK_PF_NB 256
__global__ void __launch_bounds__(256) k_centroid_split_plan(const float* __restrict__ csum, const int64_t* __restrict__ goff, int32_t G, int32_t Dp,
                                                             __nv_bfloat16* __restrict__ out, int64_t* __restrict__ goff2,
                                                             int32_t* __restrict__ col_group, int32_t* __restrict__ blockT,
                                                             int64_t* __restrict__ step0, PaGroup* __restrict__ grp, int32_t* __restrict__ col_last) {
    const int lane = threadIdx.x & 31;
    const int col = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);       // one warp per accumulator column
    const int n_blocks = (G + SDK_PF_NB - 1) / SDK_PF_NB;
    if (col >= n_blocks * SDK_PF_NB) return;
    const int b = col / SDK_PF_NB, j = col % SDK_PF_NB;
    const int g = col < G ? col : -1;
    const int64_t r0 = ((int64_t)b * 2) * SDK_PF_NB + j, r1 = r0 + SDK_PF_NB;
    if (lane == 0) {
        col_group[col] = g;
        if (j == 0) { blockT[b] = 2; step0[b] = 2 * (int64_t)b; if (b == n_blocks - 1) step0[n_blocks] = 2 * (int64_t)n_blocks; }
        if (g >= 0) {
            goff2[g] = 2 * (int64_t)g;
            if (g == G - 1) goff2[G] = 2 * (int64_t)G;
            PaGroup r;
            r.goff0 = 2 * (int64_t)g;
            r.base = (int32_t)r0;
            r.c = 1;
            grp[g] = r;
            col_last[g] = col;
        }
    }
    float rn = 0.f;
    if (g >= 0) {
        const long long n = goff[g + 1] - goff[g];
        rn = n > 0 ? 1.0f / (float)n : 0.f;
    }
    for (int e = lane; e < Dp; e += 32) {
        const float c = g >= 0 ? csum[(int64_t)g * Dp + e] * rn : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(c);
        const __nv_bfloat16 lo = __float2bfloat16_rn(c - __bfloat162float(hi));
        out[r0 * Dp + e] = __float2bfloat16_rn(2.0f * __bfloat162float(hi));
        out[r1 * Dp + e] = __float2bfloat16_rn(2.0f * __bfloat162float(lo));
    }
}

int sdk_poolfirst_applicable(const void* d_x, int32_t in_dtype, int32_t D, int32_t Dp) {
    return D % 4 == 0 && Dp % 4 == 0 && D <= 2048 && (uintptr_t)d_x % (in_dtype == SDK_IN_F16 ? 8 : 16) == 0;
}

template <typename TIn>
static int sdk_launch_normalize_centroid_t(sdk_ctx* c, const TIn* d_x, const int32_t* d_lab, int32_t label_base, int64_t n, int32_t D, int32_t Dp,
                                           float* d_f32, __nv_bfloat16* d_bf16, float* d_csum, int32_t round_bf16) {
    const int nq = (D / 4 + 31) / 32;
    const int64_t chunks = (n + SDK_PF_CHUNK - 1) / SDK_PF_CHUNK;
    int64_t blocks64 = (chunks + 7) / 8;
    int blocks = (int)(blocks64 < (int64_t)c->sm_count * 8 ? blocks64 : (int64_t)c->sm_count * 8);
#define SDK_PF_CASE(NQ) k_normalize_centroid<NQ, TIn><<<blocks, 256, 0, c->stream>>>(d_x, d_lab, label_base, n, D, Dp, d_f32, d_bf16, d_csum, round_bf16)
    if (nq <= 1) SDK_PF_CASE(1);
    else if (nq <= 2) SDK_PF_CASE(2);
    else if (nq <= 4) SDK_PF_CASE(4);
    else if (nq <= 8) SDK_PF_CASE(8);
    else SDK_PF_CASE(16);
#undef SDK_PF_CASE
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

// K1 + centroids + hi/lo split.  d_bf16 [n, Dp] = normalised operands (stage B); the pseudo-segments of the centroids go
// into c->cent_seg in the accumulate-pooling layout with their plan in c->pa_* (see k_centroid_split_plan); returns the
// number of rows of that matrix.
int sdk_launch_normalize_centroid(sdk_ctx* c, const void* d_x, int32_t in_dtype, const int32_t* d_lab, int32_t label_base, int64_t n,
                                  int32_t D, int32_t Dp, float* d_f32, __nv_bfloat16* d_bf16, const int64_t* d_goff, int32_t G,
                                  int32_t round_bf16, int64_t* n_rows_out) {
    if (!sdk_poolfirst_applicable(d_x, in_dtype, D, Dp)) return sdk_fail(c, SDK_EINVAL, "pool-first stage A needs D % 4 == 0 and an aligned segment matrix");
    const int32_t n_blocks = (G + SDK_PF_NB - 1) / SDK_PF_NB;
    const int64_t n_rows = (int64_t)n_blocks * 2 * SDK_PF_NB;
    SDK_TRY(sdk_reserve(c, c->cent_sum, (size_t)G * Dp * 4));
    SDK_TRY(sdk_reserve(c, c->cent_seg, (size_t)n_rows * Dp * 2));
    SDK_TRY(sdk_reserve(c, c->goff2, (size_t)(G + 1) * 8));
    SDK_TRY(sdk_reserve(c, c->pa_col_group, (size_t)n_blocks * SDK_PF_NB * 4));
    SDK_TRY(sdk_reserve(c, c->pa_blockT, (size_t)n_blocks * 4));
    SDK_TRY(sdk_reserve(c, c->pa_step0, (size_t)(n_blocks + 1) * 8));
    SDK_TRY(sdk_reserve(c, c->pa_grp, (size_t)G * sizeof(PaGroup)));
    SDK_TRY(sdk_reserve(c, c->pa_col_last, (size_t)G * 4));
    sdk_prof_scope ps(c, "normalize");
    float* d_csum = (float*)c->cent_sum.p;
    SDK_CUDA(c, cudaMemsetAsync(d_csum, 0, (size_t)G * Dp * 4, c->stream));
    if (n > 0) {
        if (in_dtype == SDK_IN_F16)
            SDK_TRY(sdk_launch_normalize_centroid_t<__half>(c, (const __half*)d_x, d_lab, label_base, n, D, Dp, d_f32, d_bf16, d_csum, round_bf16));
        else
            SDK_TRY(sdk_launch_normalize_centroid_t<float>(c, (const float*)d_x, d_lab, label_base, n, D, Dp, d_f32, d_bf16, d_csum, round_bf16));
    }
    k_centroid_split_plan<<<(unsigned)(((int64_t)n_blocks * SDK_PF_NB + 7) / 8), 256, 0, c->stream>>>(
        d_csum, d_goff, G, Dp, (__nv_bfloat16*)c->cent_seg.p, (int64_t*)c->goff2.p, (int32_t*)c->pa_col_group.p, (int32_t*)c->pa_blockT.p,
        (int64_t*)c->pa_step0.p, (PaGroup*)c->pa_grp.p, (int32_t*)c->pa_col_last.p);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    c->pa_blocks = n_blocks;
    c->pa_split = false;
    c->pa_chain_max = 2;
    if (n_rows_out) *n_rows_out = n_rows;
    return SDK_OK;
}

template <typename TIn>
static int sdk_launch_normalize_t(sdk_ctx* c, const TIn* d_x, int64_t n, int32_t D, int32_t Dp, float* d_f32, __nv_bfloat16* d_bf16) {
    if (n <= 0) return SDK_OK;
    sdk_prof_scope ps(c, "normalize");
    bool vec = (D % 4 == 0) && (Dp % 4 == 0) && D <= 2048 && ((uintptr_t)d_x % sdk_in<TIn>::align == 0);
    if (vec) {
        int nq = (D / 4 + 31) / 32;
        const int R = nq <= 2 ? 4 : (nq <= 4 ? 2 : 1);
        int64_t blocks64 = ((n + R - 1) / R + 7) / 8;
        int blocks = (int)(blocks64 < (int64_t)c->sm_count * 8 ? blocks64 : (int64_t)c->sm_count * 8);
#define SDK_NORM_CASE(NQ) k_normalize_vec<NQ, TIn><<<blocks, 256, 0, c->stream>>>(d_x, n, D, Dp, d_f32, d_bf16)
        if (nq <= 1) SDK_NORM_CASE(1);
        else if (nq <= 2) SDK_NORM_CASE(2);
        else if (nq <= 4) SDK_NORM_CASE(4);
        else if (nq <= 8) SDK_NORM_CASE(8);
        else SDK_NORM_CASE(16);
#undef SDK_NORM_CASE
    } else {
        int64_t blocks64 = (n + 7) / 8;
        int blocks = (int)(blocks64 < (int64_t)c->sm_count * 8 ? blocks64 : (int64_t)c->sm_count * 8);
        k_normalize_generic<TIn><<<blocks, 256, 0, c->stream>>>(d_x, n, D, Dp, d_f32, d_bf16);
    }
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

int sdk_launch_normalize_in(sdk_ctx* c, const void* d_x, int32_t in_dtype, int64_t n, int32_t D, int32_t Dp, float* d_f32,
                            __nv_bfloat16* d_bf16) {
    if (in_dtype == SDK_IN_F16) return sdk_launch_normalize_t<__half>(c, (const __half*)d_x, n, D, Dp, d_f32, d_bf16);
    return sdk_launch_normalize_t<float>(c, (const float*)d_x, n, D, Dp, d_f32, d_bf16);
}
int sdk_launch_normalize(sdk_ctx* c, const float* d_x, int64_t n, int32_t D, int32_t Dp, float* d_f32,
                         __nv_bfloat16* d_bf16) {
    return sdk_launch_normalize_t<float>(c, d_x, n, D, Dp, d_f32, d_bf16);
}

// goff[g] = first segment index whose (label - label_base) >= g  (labels non-decreasing, in [base, base+L))
__global__ void k_group_offsets(const int32_t* __restrict__ lab, int64_t N, int32_t L, int32_t label_base,
                                int64_t* __restrict__ goff, int32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (N == 0) {
        if (i <= L) goff[i] = 0;
        return;
    }
    if (i > N) return;
    int32_t a = (i == 0) ? -1 : lab[i - 1] - label_base;
    int32_t b = (i == N) ? L : lab[i] - label_base;
    if (i < N && (b < 0 || b >= L)) { atomicOr(flag, 1); return; }
    if (b < a) { atomicOr(flag, 2); return; }
    for (int32_t g = a + 1; g <= b; ++g) goff[g] = i;
}

int sdk_launch_group_offsets(sdk_ctx* c, const int32_t* d_lab, int64_t N, int32_t L, int32_t label_base,
                             int64_t* d_goff, int32_t* d_flag) {
    int64_t work = N == 0 ? (int64_t)L + 1 : N + 1;
    int blocks = (int)((work + 255) / 256);
    k_group_offsets<<<blocks, 256, 0, c->stream>>>(d_lab, N, L, label_base, d_goff, d_flag);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}
