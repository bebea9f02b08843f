// select.cu -- K3 (ordered top-k with row->speaker max and threshold), the top-k certificate of the
// tcgen05 path, the fp64 assignment (combine_signals restatement) and K4 (merge after all-gather).
#include "common.cuh"

#define SDK_SEL_THREADS 256

__device__ __forceinline__ unsigned long long sdk_block_max_u64(unsigned long long v, unsigned long long* sh) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o > v ? o : v;
    }
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[w] = v;
    __syncthreads();
    unsigned long long r = sh[0];
    for (int i = 1; i < nw; ++i) r = sh[i] > r ? sh[i] : r;
    return r;
}

// One CTA per label group.  Entries are (slot -> bank row, pooled Q30).  Keys are
// (orderable fp32 score << 32) | ~row, so a descending key order is exactly (-score, row).
// Keys are unique, so "already extracted" == key >= last extracted key: no per-entry state.
// A speaker's first extracted row is its best row (max over rows, ties: lowest row); later rows of
// the same speaker are skipped.  Stops at k matches or when the score drops below the threshold.
__global__ void __launch_bounds__(SDK_SEL_THREADS)
k_select(long long* __restrict__ qpool, const int64_t* __restrict__ goff, const int32_t* __restrict__ glist,
         const int32_t* __restrict__ cand_row, int64_t nslot, int32_t pool,
         const int32_t* __restrict__ row_speaker, const uint8_t* __restrict__ row_trust, double threshold,
         int32_t k, int64_t row_offset, const float* __restrict__ gbound, float eps_base, float eps_chain, const PaGroup* __restrict__ grp,
         int32_t chain_div, int32_t* __restrict__ fb_count, int32_t* __restrict__ fb_list, int64_t* __restrict__ out_row,
         float* __restrict__ out_score, int32_t* __restrict__ out_count, uint8_t* __restrict__ out_trust,
         int32_t* __restrict__ out_spk) {
    __shared__ unsigned long long sh[SDK_SEL_THREADS / 32];
    __shared__ int32_t sel_spk[SDK_MAX_K];
    const int32_t gi = blockIdx.x;
    const int32_t g = glist ? glist[gi] : gi;
    const int tid = threadIdx.x;
    const long long n = goff[g + 1] - goff[g];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(qpool) + (int64_t)gi * nslot;

    for (int i = tid; i < k; i += SDK_SEL_THREADS) {
        out_row[(int64_t)g * k + i] = -1;
        out_score[(int64_t)g * k + i] = 0.f;
        out_trust[(int64_t)g * k + i] = SDK_TRUST_UNKNOWN;
        out_spk[(int64_t)g * k + i] = -1;
    }
    if (n <= 0) {
        if (tid == 0) out_count[g] = 0;
        return;
    }
    // pass 0: Q30 -> key, in place (0 = invalid slot)
    for (int64_t j = tid; j < nslot; j += SDK_SEL_THREADS) {
        int32_t row = cand_row ? cand_row[(int64_t)gi * nslot + j] : (int32_t)j;
        unsigned long long key = 0ull;
        if (row >= 0) {
            float sim = sdk_pool_finish((long long)keys[j], n, pool);
            key = ((unsigned long long)sdk_fkey(sim) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)row);
        }
        keys[j] = key;
    }
    __syncthreads();
    unsigned long long last = ~0ull;
    int cnt = 0;
    float kth = 0.f;
    while (cnt < k) {
        unsigned long long best = 0ull;
        for (int64_t j = tid; j < nslot; j += SDK_SEL_THREADS) {
            unsigned long long key = keys[j];
            if (key < last && key > best) best = key;
        }
        best = sdk_block_max_u64(best, sh);
        if (best == 0ull) break;
        last = best;
        float sim = sdk_funkey((uint32_t)(best >> 32));
        if (!((double)sim >= threshold)) break;
        int32_t row = (int32_t)(0xffffffffu - (uint32_t)(best & 0xffffffffu));
        int32_t spk = row_speaker[row];
        bool dup = false;
        for (int i = 0; i < cnt; ++i) dup |= sel_spk[i] == spk;
        __syncthreads();
        if (dup) continue;
        if (tid == 0) {
            sel_spk[cnt] = spk;
            out_row[(int64_t)g * k + cnt] = (int64_t)row + row_offset;
            out_score[(int64_t)g * k + cnt] = sim;
            out_trust[(int64_t)g * k + cnt] = row_trust ? row_trust[row] : SDK_TRUST_UNKNOWN;
            out_spk[(int64_t)g * k + cnt] = spk;
        }
        kth = sim;
        ++cnt;
        __syncthreads();
    }
    if (tid == 0) {
        out_count[g] = cnt;
        if (gbound) {
            // certificate: every row that was NOT re-scored has approx score <= bound, hence a
            // canonical score <= bound + eps.  It cannot enter the result if that is below the
            // threshold, or below the k-th kept score when the list is full.
            // eps_g bounds |stage-A pooled score - canonical| for THIS group: eps_base covers one segment's dot product (and,
            // for fp32 banks, the bf16 rounding of the stage-A operands); eps_chain scales with the group's accumulation
            // chain -- segments per accumulator column when pooling happens inside the MMA accumulation (grp != null:
            // the partial sum grows with every segment, and so does the rounding unit), blocks of 32 columns when the
            // epilogue pools (chain_div == 32), nothing for max pooling.  Model and derivation: DESIGN.md section 2.
            const double chain = grp ? (double)((n + grp[g].c - 1) / grp[g].c) : (chain_div > 0 ? (double)(n / chain_div + 70) : 0.0);
            const double eps_g = (double)eps_base + (double)eps_chain * chain;
            double b = (double)gbound[g] + eps_g;
            bool safe = (b < threshold) || (cnt == k && b < (double)kth);
            if (!safe) {
                int pos = atomicAdd(fb_count, 1);
                fb_list[pos] = g;
            }
        }
    }
}

int sdk_launch_select(sdk_ctx* c, const long long* d_qpool, const int64_t* d_goff, const int32_t* d_glist,
                      int32_t ngroups, const int32_t* d_cand_row, int64_t nslot, int32_t pool,
                      const int32_t* d_row_speaker, const uint8_t* d_row_trust, double threshold, int32_t k,
                      int64_t row_offset, const float* d_gbound, float eps_base, float eps_chain, const PaGroup* d_grp, int32_t chain_div,
                      int32_t* d_fb_count, int32_t* d_fb_list, int64_t* d_out_row, float* d_out_score, int32_t* d_out_count,
                      uint8_t* d_out_trust, int32_t* d_out_spk) {
    if (ngroups <= 0) return SDK_OK;
    sdk_prof_scope ps(c, "select");
    k_select<<<ngroups, SDK_SEL_THREADS, 0, c->stream>>>(const_cast<long long*>(d_qpool), d_goff, d_glist, d_cand_row,
                                                         nslot, pool, d_row_speaker, d_row_trust, threshold, k,
                                                         row_offset, d_gbound, eps_base, eps_chain, d_grp, chain_div, d_fb_count, d_fb_list, d_out_row,
                                                         d_out_score, d_out_count, d_out_trust, d_out_spk);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

// ---- assignment: restates speaker-assign:418-492 for embedding_match-only signal lists ---------
// (see oracle/canonical.c orc_assign for the line-by-line citation).  One WARP per label group, lane = match; fp64, same
// operations as the Python: weight = 0.4; weight *= mult; ws = weight * score; scores[id] = 0.0 + ws; stable descending
// sort (= rank by (ws desc, position asc): the insertion sort of the restatement moves an entry only past strictly smaller
// ones); bands; threshold.  Every lane loads its match at once, so a label costs one memory round trip instead of one per
// match (the one-thread-per-label form measured 9.6 us for 8 labels).
__global__ void __launch_bounds__(128)
k_assign(const int64_t* __restrict__ m_row, const float* __restrict__ m_score,
         const uint8_t* __restrict__ m_trust, const int32_t* __restrict__ m_count, int32_t L,
         int32_t k, double thr, int32_t min_trust, int32_t* __restrict__ a_idx,
         double* __restrict__ a_score, int32_t* __restrict__ a_conf, int32_t* __restrict__ c_idx,
         double* __restrict__ c_score) {
    const int32_t g = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    // launched as a programmatic dependent (its launch latency hides under the kernel before it): wait for that kernel's
    // results; a no-op after a kernel that never signals early
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (g >= L) return;
    auto rank_of = [](int code) { return code == 2 ? 0 : code == 1 ? 1 : code == 0 ? 2 : -1; };
    const int min_rank = rank_of(min_trust);
    const int cnt = m_count[g];
    bool live = false;
    double wa = 0.0;
    if (lane < cnt && lane < k) {
        const int64_t row = m_row[(int64_t)g * k + lane];
        const int t = m_trust[(int64_t)g * k + lane];
        const double sc = (double)m_score[(int64_t)g * k + lane];
        const int tr = rank_of(t);
        live = row >= 0 && !(min_rank >= 0 && tr >= 0 && tr < min_rank);
        const double mult = t == 0 ? 1.0 : t == 1 ? 0.7 : t == 2 ? 0.4 : t == 3 ? 0.0 : 0.5;
        double weight = 0.4;
        weight = __dmul_rn(weight, mult);
        wa = __dadd_rn(0.0, __dmul_rn(weight, sc));
    }
    const uint32_t mlive = __ballot_sync(0xffffffffu, live);
    const int m = __popc(mlive);
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const double o = __shfl_sync(0xffffffffu, wa, j);
        if ((mlive >> j) & 1u) rank += (o > wa || (o == wa && j < lane)) ? 1 : 0;
    }
    if (!live) rank = 64;
    // lanes holding ranks 0..3
    int src[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { const uint32_t mr = __ballot_sync(0xffffffffu, rank == r); src[r] = mr ? __ffs(mr) - 1 : -1; }
    double w4[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) w4[r] = __shfl_sync(0xffffffffu, wa, src[r] < 0 ? 0 : src[r]);
    if (lane != 0) return;
    for (int j = 0; j < 3; ++j) { c_idx[g * 3 + j] = -1; c_score[g * 3 + j] = 0.0; }
    if (m == 0) { a_idx[g] = -1; a_score[g] = 0.0; a_conf[g] = SDK_CONF_UNASSIGNED; return; }
    const double best = w4[0];
    const int conf = best >= 0.7 ? SDK_CONF_HIGH : best >= 0.4 ? SDK_CONF_MEDIUM : best >= 0.2 ? SDK_CONF_LOW : SDK_CONF_UNASSIGNED;
    a_score[g] = best;
    if (best < thr) {
        a_idx[g] = -1;
        a_conf[g] = SDK_CONF_UNASSIGNED;
        for (int j = 0; j < 3 && j < m; ++j) { c_idx[g * 3 + j] = src[j]; c_score[g * 3 + j] = w4[j]; }
    } else {
        a_idx[g] = src[0];
        a_conf[g] = conf;
        for (int j = 0; j < 3 && j + 1 < m; ++j) { c_idx[g * 3 + j] = src[j + 1]; c_score[g * 3 + j] = w4[j + 1]; }
    }
}

int sdk_launch_assign(sdk_ctx* c, const int64_t* d_row, const float* d_score, const uint8_t* d_trust,
                      const int32_t* d_count, int32_t L, int32_t k, double thr, int32_t min_trust, int32_t* d_idx,
                      double* d_ascore, int32_t* d_conf, int32_t* d_cidx, double* d_cscore) {
    if (L <= 0) return SDK_OK;
    sdk_prof_scope ps(c, "assign");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((L + 3) / 4));
    cfg.blockDim = dim3(128);
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SDK_CUDA(c, cudaLaunchKernelEx(&cfg, k_assign, d_row, d_score, d_trust, d_count, L, k, thr, min_trust, d_idx, d_ascore, d_conf, d_cidx, d_cscore));
    c->launches++;
    return SDK_OK;
}

// ---- K4: merge `world` per-rank top-k lists (after ncclAllGather) by (-score, global row) -------
// Bank shards are cut on speaker boundaries (asserted by the host), so no cross-rank speaker
// de-duplication is needed.  One warp per label group; world*k <= 8*32 entries, 8 per lane.
// The lists are read in place from the gathered result records (rank r's record starts at all + r*stride and has the
// layout of the local record: offsets o_*), so the all-gather needs no packing or unpacking copies.
struct MergeSrc {
    const char* all;
    size_t stride, off_row, off_score, off_spk, off_count, off_trust;
};
__global__ void k_merge_topk(const MergeSrc src, int32_t world, int32_t L, int32_t k,
                             int64_t* __restrict__ o_row, float* __restrict__ o_score,
                             int32_t* __restrict__ o_count, uint8_t* __restrict__ o_trust,
                             int32_t* __restrict__ o_spk, int32_t* __restrict__ o_flags) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) {
        // a peer that failed locally (status word) or saw bad labels poisons the merged result: tell the host at fetch time
        int bad = 0x7fffffff;
        for (int r = lane; r < world; r += 32) {
            const int32_t* f = reinterpret_cast<const int32_t*>(src.all + (size_t)r * src.stride);
            if (f[SDK_FLAG_STATUS] != 0 || f[SDK_FLAG_LABEL] != 0) bad = r < bad ? r : bad;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, bad, off); bad = o < bad ? o : bad; }
        if (lane == 0 && bad != 0x7fffffff) o_flags[SDK_FLAG_PEER] = bad + 1;
    }
    if (warp >= L) return;
    const int g = warp;
    const int total = world * k;
    int taken = 0;
    uint32_t last_key = 0;
    long long last_row = -1;
    for (int it = 0; it < k; ++it) {
        // best entry strictly after (last score, last row) in (-score, row) order
        uint32_t bkey = 0;
        long long brow = 0x7fffffffffffffffLL;
        int bidx = -1;                                    // rank * k + index
        for (int e = lane; e < total; e += 32) {
            const int r = e / k, i = e - r * k;
            const char* rec = src.all + (size_t)r * src.stride;
            if (i >= reinterpret_cast<const int32_t*>(rec + src.off_count)[g]) continue;
            const int64_t off = (int64_t)g * k + i;
            const long long row = reinterpret_cast<const int64_t*>(rec + src.off_row)[off];
            if (row < 0) continue;
            const uint32_t key = sdk_fkey(reinterpret_cast<const float*>(rec + src.off_score)[off]);
            const bool after = (it == 0) || key < last_key || (key == last_key && row > last_row);
            if (!after) continue;
            if (bidx < 0 || key > bkey || (key == bkey && row < brow)) { bkey = key; brow = row; bidx = e; }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            uint32_t ok = __shfl_xor_sync(0xffffffffu, bkey, off);
            long long orow = __shfl_xor_sync(0xffffffffu, brow, off);
            int oidx = __shfl_xor_sync(0xffffffffu, bidx, off);
            if (oidx >= 0 && (bidx < 0 || ok > bkey || (ok == bkey && orow < brow))) { bkey = ok; brow = orow; bidx = oidx; }
        }
        if (bidx < 0) break;
        if (lane == 0) {
            const int r = bidx / k, i = bidx - r * k;
            const char* rec = src.all + (size_t)r * src.stride;
            const int64_t off = (int64_t)g * k + i;
            o_row[(int64_t)g * k + taken] = brow;
            o_score[(int64_t)g * k + taken] = reinterpret_cast<const float*>(rec + src.off_score)[off];
            o_trust[(int64_t)g * k + taken] = reinterpret_cast<const uint8_t*>(rec + src.off_trust)[off];
            o_spk[(int64_t)g * k + taken] = reinterpret_cast<const int32_t*>(rec + src.off_spk)[off];
        }
        last_key = bkey;
        last_row = brow;
        ++taken;
    }
    if (lane == 0) {
        o_count[g] = taken;
        for (int i = taken; i < k; ++i) {
            o_row[(int64_t)g * k + i] = -1;
            o_score[(int64_t)g * k + i] = 0.f;
            o_trust[(int64_t)g * k + i] = SDK_TRUST_UNKNOWN;
            o_spk[(int64_t)g * k + i] = -1;
        }
    }
}

int sdk_launch_merge_topk(sdk_ctx* c, const void* d_all, size_t stride, const sdk_out_view& v, int32_t world, int32_t L, int32_t k) {
    if (L <= 0) return SDK_OK;
    sdk_prof_scope ps(c, "merge_topk");
    MergeSrc src;
    src.all = (const char*)d_all;
    src.stride = stride;
    src.off_row = v.off_row; src.off_score = v.off_score; src.off_spk = v.off_spk; src.off_count = v.off_count; src.off_trust = v.off_trust;
    int threads = 128, warps_per_block = threads / 32;
    int blocks = (L + warps_per_block - 1) / warps_per_block;
    k_merge_topk<<<blocks, threads, 0, c->stream>>>(src, world, L, k, v.row, v.score, v.count, v.trust, v.spk, v.flags);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

// empty lists for L label groups (a rank with no bank rows, or one that failed before the collective)
__global__ void k_fill_empty(int64_t* __restrict__ row, float* __restrict__ score, int32_t* __restrict__ spk, int32_t* __restrict__ count,
                             uint8_t* __restrict__ trust, int32_t L, int32_t k) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)L * k) { row[i] = -1; score[i] = 0.f; spk[i] = -1; trust[i] = SDK_TRUST_UNKNOWN; }
    if (i < L) count[i] = 0;
}
int sdk_launch_fill_empty(sdk_ctx* c, const sdk_out_view& v, int32_t L, int32_t k) {
    if (L <= 0) return SDK_OK;
    const int64_t n = (int64_t)L * k;
    k_fill_empty<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(v.row, v.score, v.spk, v.count, v.trust, L, k);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}
