// poolacc.cu -- K2c: mean pooling fused INTO the tensor-core accumulation ("accumulate-pooling").
//
// poolgemm.cu puts segments on the accumulator columns and pools them in the epilogue: every scored pair is read
// out of TMEM once and added on the CUDA cores.  With D <= 256 a 128x256 tile is at most 16 MMAs (2048 tensor cycles)
// but 128 KB of TMEM read-out + 32k FADDs, so the epilogue, not the tensor pipe, is the bound (profiles/r01_*).
//
// For MEAN pooling the sum over a label's segments can be done by the MMA itself.  Segments are laid out
// "group interleaved": every label group owns c >= 1 accumulator COLUMNS of a block of 256 columns and deals its
// segments over them round robin; step t of a block is the 256-row slab holding the t-th segment of each of its
// columns (zero rows once a column is exhausted).  Issuing the MMAs of step 0,1,..,T-1 into the SAME TMEM tile
// (accumulate = true across steps as well as across K) leaves
//        D[row, j] = sum_t  bank[row] . seg_t(column j)           -- a partial pooled sum, every pair contracted --
// after T*D/16 MMAs; the epilogue adds the c columns of a group and scales by 1/n.  The tile is read out once per
// T*256 segments instead of once per 256 segments, the B ring needs no chunk residency, and the kernel behaves like a
// large-K GEMM.
//
// Plans (built on the device per call).  (A) many groups (> 2048): c = 1, groups sorted by size (counting sort), blocks of
// 256 sorted groups -- hour-long recordings by the ten thousand (config 3).  (B1) at most 256 groups: the groups are dealt
// into nb bins (= blocks) by LPT, every block gets its own step count T_b = min T with sum max(1, ceil(n / T)) <= 256 so
// that every block is full, and nb is chosen by simulating the kernel's static unit -> CTA schedule (wave quantisation
// included) -- the 16-label affinity of config 5 and single-meeting shapes.  (B2) 257..2048 groups: next-fit packing of
// c = ceil(n / T) columns per group for ~60 candidate step caps T under the same cost model.  No group straddles a block.
//
// Kernel pipeline: TMA producer warp / MMA-issuing warp / epilogue warpgroups.  The bank tiles of a unit are handed over per
// K chunk (the next unit's chunk kc is loaded as soon as the last step of this unit has consumed chunk kc); with one row
// tile per unit (MT == 1) two accumulator sets alternate in TMEM so that the read-out of unit u overlaps the MMAs of u + 1.
//
// Everything downstream is unchanged: the epilogue flushes per (group, 32 bank rows) candidate slots, k_pg_merge picks
// the candidates, exact.cu re-scores them canonically (reading segments from the interleaved layout), select.cu
// certifies.  Max pooling stays on poolgemm.cu.
#include <stdlib.h>

#include "tcgen05.cuh"

#define PA_NB 256                 // accumulator columns per block
#define PA_BUCKETS 65536
#define PA_SMALL_G 2048           // plan (B) up to this many label groups
#define PA_NCAND 64               // candidate step caps evaluated by plan (B)

// ---- plan (A): counting sort of the groups by (clipped) size, descending; one column per group ----------------------
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
// per block of 256 sorted groups: T_b = largest group; column -> group table (-1 = padding column)
__global__ void k_pa_blocks(const int64_t* __restrict__ goff, const int32_t* __restrict__ sorted_group, int32_t G, int32_t n_blocks,
                            int32_t* __restrict__ col_group, int32_t* __restrict__ blockT) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += gridDim.x * blockDim.x) {
        int64_t m = 0;
        for (int j = 0; j < PA_NB; ++j) {
            int pos = b * PA_NB + j;
            int g = pos < G ? sorted_group[pos] : -1;
            col_group[pos] = g;
            if (g >= 0) { int64_t n = goff[g + 1] - goff[g]; m = n > m ? n : m; }
        }
        blockT[b] = (int32_t)(m > 0x7fffffff ? 0x7fffffff : m);
    }
}
// step0 = exclusive prefix of T_b (one warp); plan_out = {total steps, blocks}
__global__ void k_pa_steps(const int32_t* __restrict__ blockT, int32_t n_blocks, int64_t* __restrict__ step0, int64_t* __restrict__ plan_out) {
    const int lane = threadIdx.x;
    int64_t carry = 0, tmax = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 32) {
        const int b = b0 + lane;
        int64_t v = b < n_blocks ? blockT[b] : 0, incl = v;
        tmax = v > tmax ? v : tmax;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int64_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        if (b < n_blocks) step0[b] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { const int64_t o = __shfl_xor_sync(0xffffffffu, tmax, off); tmax = o > tmax ? o : tmax; }
    if (lane == 0) { step0[n_blocks] = carry; plan_out[0] = carry; plan_out[1] = n_blocks; plan_out[2] = tmax; }
}
__global__ void k_pa_group_rows(const int64_t* __restrict__ goff, const int32_t* __restrict__ group_pos, const int64_t* __restrict__ step0,
                                int32_t G, PaGroup* __restrict__ grp, int32_t* __restrict__ col_last) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int pos = group_pos[g];
    PaGroup r;
    r.goff0 = goff[g];
    const int64_t base = step0[pos / PA_NB] * PA_NB + (pos % PA_NB);
    r.base = (int32_t)(base > 0x7fffffff ? 0x7fffffff : base);      // the host rejects layouts of more than 2^31-1 rows
    r.c = 1;
    grp[g] = r;
    col_last[g] = pos;
}

// ---- plan (B): few groups, columns split; one CTA ------------------------------------------------------------------
struct PaCost {
    float step_cycles;      // tensor cycles of one 256-row step of a unit (KCH * 4 * MT MMAs of 128 x 256 x 16)
    float unit_cycles;      // exposed per unit: bank tile load + TMEM read-out
    int32_t RB, sms, t_min;
};
__device__ __forceinline__ int pa_cand_T(int i, int t_min) {      // geometric ladder, ratio 2^(1/4)
    float t = (float)t_min * exp2f(0.25f * (float)i);
    int T = (int)(t + 0.5f);
    return T > 65535 ? 65535 : T;
}
__global__ void __launch_bounds__(1024)
k_pa_plan_small(const int64_t* __restrict__ goff, int32_t G, PaCost pc, int32_t max_blocks, int32_t* __restrict__ col_group,
                int32_t* __restrict__ col_meta, int32_t* __restrict__ blockT, int64_t* __restrict__ step0, int64_t* __restrict__ plan_out,
                PaGroup* __restrict__ grp, int32_t* __restrict__ col_last) {
    __shared__ int32_t sn[PA_SMALL_G];          // group sizes
    __shared__ int16_t sorder[PA_SMALL_G];      // groups by size, descending (ties: lower id first)
    __shared__ int32_t scol0[PA_SMALL_G];       // first column of the group (global column index)
    __shared__ int16_t scnt[PA_SMALL_G];        // columns of the group
    __shared__ float scost[PA_NCAND];
    __shared__ int sbest, s_nblocks;
    const int tid = threadIdx.x;
    for (int g = tid; g < G; g += blockDim.x) {
        int64_t n = goff[g + 1] - goff[g];
        sn[g] = (int32_t)(n > 0x7fffffff ? 0x7fffffff : n);
    }
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x) {
        const int32_t n = sn[g];
        int rank = 0;
        for (int j = 0; j < G; ++j) { const int32_t m = sn[j]; rank += (m > n || (m == n && j < g)) ? 1 : 0; }
        sorder[rank] = (int16_t)g;
    }
    __syncthreads();
    // every candidate step cap is simulated exactly (next-fit in size order) by one thread
    if (tid < PA_NCAND) {
        const int T = pa_cand_T(tid, pc.t_min);
        int fill = 0, Tb = 0, nb = 0, Tmax = 0;
        long long steps = 0;
        for (int i = 0; i < G; ++i) {
            const int32_t n = sn[sorder[i]];
            int c = n > 0 ? (n + T - 1) / T : 1;
            if (c > PA_NB) c = PA_NB;
            const int s = n > 0 ? (n + c - 1) / c : 0;
            if (fill + c > PA_NB) { steps += Tb; ++nb; Tmax = Tb > Tmax ? Tb : Tmax; fill = 0; Tb = 0; }
            fill += c;
            Tb = s > Tb ? s : Tb;
        }
        if (fill > 0) { steps += Tb; ++nb; Tmax = Tb > Tmax ? Tb : Tmax; }
        const float units = (float)nb * (float)pc.RB;
        const float waves = ceilf(units / (float)pc.sms);
        // MMA time: perfectly spread steps, plus one longest unit of tail when the units do not fill the last wave
        float cost = ((float)steps * (float)pc.RB / (float)pc.sms) * pc.step_cycles + waves * pc.unit_cycles;
        const float frac = units / (float)pc.sms - floorf(units / (float)pc.sms);
        if (units < (float)pc.sms) cost = (float)Tmax * pc.step_cycles + pc.unit_cycles;     // single partial wave
        else if (frac > 0.f) cost += (1.f - frac) * (float)Tmax * pc.step_cycles * 0.5f;
        if (nb > max_blocks) cost = 3.0e38f;
        scost[tid] = cost;
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int i = 1; i < PA_NCAND; ++i) if (scost[i] < scost[best]) best = i;
        if (!(scost[best] < 3.0e38f)) best = PA_NCAND - 1;      // cannot happen: the largest cap needs the fewest blocks
        sbest = best;
        const int T = pa_cand_T(best, pc.t_min);
        int fill = 0, Tb = 0, nb = 0, Tlong = 0;
        long long steps = 0;
        for (int i = 0; i < G; ++i) {
            const int g = sorder[i];
            const int32_t n = sn[g];
            int c = n > 0 ? (n + T - 1) / T : 1;
            if (c > PA_NB) c = PA_NB;
            const int s = n > 0 ? (n + c - 1) / c : 0;
            Tlong = s > Tlong ? s : Tlong;
            if (fill + c > PA_NB) { blockT[nb] = Tb; step0[nb] = steps; steps += Tb; ++nb; fill = 0; Tb = 0; }
            scol0[g] = nb * PA_NB + fill;
            scnt[g] = (int16_t)c;
            fill += c;
            Tb = s > Tb ? s : Tb;
        }
        if (fill > 0) { blockT[nb] = Tb; step0[nb] = steps; steps += Tb; ++nb; }
        step0[nb] = steps;
        plan_out[0] = steps;
        plan_out[1] = nb;
        plan_out[2] = Tlong;
        s_nblocks = nb;
    }
    __syncthreads();
    const int ncol = s_nblocks * PA_NB;
    for (int i = tid; i < ncol; i += blockDim.x) { col_group[i] = -1; col_meta[i] = -1; }
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x) {
        const int c0 = scol0[g], c = scnt[g];
        for (int j = 0; j < c; ++j) { col_group[c0 + j] = g; col_meta[c0 + j] = j == c - 1 ? g : -2; }
        PaGroup r;
        r.goff0 = goff[g];
        const int64_t base = step0[c0 / PA_NB] * PA_NB + (c0 % PA_NB);
        r.base = (int32_t)(base > 0x7fffffff ? 0x7fffffff : base);
        r.c = c;
        grp[g] = r;
        col_last[g] = c0 + c - 1;
    }
}

// ---- plan (B1): at most PA_BINS_G (256) groups -- balanced bins, one step count per block ------------------------------------
// With a handful of groups (the 16 labels of config 5, the 8 labels of one meeting) next-fit leaves blocks part empty and
// the number of units rarely divides over the SMs.  Here the groups are dealt into nb bins (= blocks) by LPT, every
// block gets ITS OWN step count T_b = min T with sum_g max(1, ceil(n_g / T)) <= 256 -- so every block is full -- and
// nb is chosen by simulating the kernel's static unit -> CTA assignment (wave quantisation included).  Warp per candidate.
#define PA_BINS_G 256
#define PA_MAX_BINS 64
#define PA_BIN_CAP 250            // label groups per bin (every group needs at least one of the 256 columns)
__device__ __forceinline__ void pa_warp_argmin(long long& v, int& idx) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const long long ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}
// one LPT pass of the size-sorted groups over nb bins; lane owns bins lane and lane + 32 (a full bin takes no more groups)
__device__ __forceinline__ void pa_lpt(const int32_t* sn, const int16_t* sorder, int G, int nb, int lane, long long& load0, long long& load1,
                                       int& cnt0, int& cnt1, int16_t* sbin /* may be null */) {
    const long long INF = 0x7fffffffffffffffLL;
    load0 = 0; load1 = 0; cnt0 = 0; cnt1 = 0;
    for (int i = 0; i < G; ++i) {
        const int g = sorder[i];
        const int32_t n = sn[g];
        const long long e0 = (lane < nb && cnt0 < PA_BIN_CAP) ? load0 : INF, e1 = (lane + 32 < nb && cnt1 < PA_BIN_CAP) ? load1 : INF;
        long long v = e0 <= e1 ? e0 : e1;
        int idx = e0 <= e1 ? lane : lane + 32;
        pa_warp_argmin(v, idx);
        if (idx == lane) { load0 += n; ++cnt0; }
        if (idx == lane + 32) { load1 += n; ++cnt1; }
        if (sbin && lane == 0) sbin[g] = (int16_t)idx;
    }
}
__global__ void __launch_bounds__(1024)
k_pa_plan_bins(const int64_t* __restrict__ goff, int32_t G, PaCost pc, int32_t* __restrict__ col_group, int32_t* __restrict__ col_meta,
               int32_t* __restrict__ blockT, int64_t* __restrict__ step0, int64_t* __restrict__ plan_out, PaGroup* __restrict__ grp,
               int32_t* __restrict__ col_last) {
    __shared__ int32_t sn[PA_BINS_G];
    __shared__ int16_t sorder[PA_BINS_G];
    __shared__ int16_t sbin[PA_BINS_G];
    __shared__ int16_t slist[PA_BINS_G];       // groups bin by bin
    __shared__ int32_t soff[PA_MAX_BINS + 1], scur[PA_MAX_BINS];
    __shared__ int32_t sT[PA_MAX_BINS];
    __shared__ int32_t sTw[32][PA_MAX_BINS];   // per warp: estimated T_b of the candidate being evaluated
    __shared__ float scost[32];
    __shared__ int s_nb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int g = tid; g < G; g += blockDim.x) {
        int64_t n = goff[g + 1] - goff[g];
        sn[g] = (int32_t)(n > 0x7fffffff ? 0x7fffffff : n);
    }
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x) {
        const int32_t n = sn[g];
        int rank = 0;
        for (int j = 0; j < G; ++j) { const int32_t m = sn[j]; rank += (m > n || (m == n && j < g)) ? 1 : 0; }
        sorder[rank] = (int16_t)g;
    }
    __syncthreads();
    // candidates: nb_min, nb_min + 1, ... one per warp
    const int nb_min = (G + PA_BIN_CAP - 1) / PA_BIN_CAP;
    const int nb_hi = G < PA_MAX_BINS ? G : PA_MAX_BINS;
    {
        const int nb = nb_min + warp;
        if (nb > nb_hi) {
            if (lane == 0) scost[warp] = 3.0e38f;
        } else {
            long long load0, load1;
            int cnt0, cnt1;
            pa_lpt(sn, sorder, G, nb, lane, load0, load1, cnt0, cnt1, nullptr);
            // estimate of T_b (sufficient, a few percent high): load / (256 - groups)
            int T0 = 0, T1 = 0;
            if (lane < nb) { const int d = 256 - cnt0 > 1 ? 256 - cnt0 : 1; T0 = (int)((load0 + d - 1) / d); }
            if (lane + 32 < nb) { const int d = 256 - cnt1 > 1 ? 256 - cnt1 : 1; T1 = (int)((load1 + d - 1) / d); }
            sTw[warp][lane] = T0;
            sTw[warp][lane + 32] = T1;
            __syncwarp();
            // the kernel's schedule: unit u = b * RB + rb runs on CTA u % sms, i.e. CTA i gets RB / sms units of block b,
            // plus one if (i - first CTA of the block) mod sms < RB % sms
            float worst = 0.f;
            const int uq = pc.RB / pc.sms, ur = pc.RB % pc.sms;
            for (int i = lane; i < pc.sms; i += 32) {
                float t = 0.f;
                int u0m = 0;                                                     // (b * RB) % sms
                for (int b = 0; b < nb; ++b) {
                    const int Tb = sTw[warp][b];
                    int d = i - u0m;
                    d += d < 0 ? pc.sms : 0;
                    const int cntu = uq + (d < ur ? 1 : 0);
                    if (Tb > 0) t += (float)cntu * ((float)Tb * pc.step_cycles + pc.unit_cycles);
                    u0m += ur;
                    u0m -= u0m >= pc.sms ? pc.sms : 0;
                }
                worst = fmaxf(worst, t);
            }
            __syncwarp();
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, off));
            if (lane == 0) scost[warp] = worst;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int i = 1; i < 32; ++i) if (scost[i] < scost[best]) best = i;
        s_nb = nb_min + best;
    }
    __syncthreads();
    const int nb = s_nb;
    if (warp == 0) {
        long long load0, load1;
        int cnt0, cnt1;
        pa_lpt(sn, sorder, G, nb, lane, load0, load1, cnt0, cnt1, sbin);
        // groups bin by bin (counting sort on the bin index)
        if (lane < PA_MAX_BINS / 2) { scur[lane] = 0; scur[lane + 32] = 0; }
        __syncwarp();
        sTw[0][lane] = lane < nb ? cnt0 : 0;
        sTw[0][lane + 32] = lane + 32 < nb ? cnt1 : 0;
        __syncwarp();
        if (lane == 0) {
            int acc = 0;
            for (int b = 0; b < nb; ++b) { soff[b] = acc; acc += sTw[0][b]; }
            soff[nb] = acc;
        }
        __syncwarp();
        for (int g = lane; g < G; g += 32) {
            const int b = sbin[g];
            slist[soff[b] + atomicAdd(&scur[b], 1)] = (int16_t)g;
        }
        __syncwarp();
        // exact T_b: smallest T with sum max(1, ceil(n / T)) <= 256 over the bin's groups
        for (int b = lane; b < nb; b += 32) {
            const int l0 = soff[b], l1 = soff[b + 1];
            long long tot = 0;
            int32_t mx = 0;
            for (int i = l0; i < l1; ++i) { const int32_t n = sn[slist[i]]; tot += n; mx = n > mx ? n : mx; }
            int lo = 1, hi = mx > 1 ? mx : 1;                       // T = max n always fits (<= 250 groups, one column each)
            if (tot == 0) { sT[b] = 0; continue; }
            while (lo < hi) {
                const int T = lo + (hi - lo) / 2;
                int cols = 0;
                for (int i = l0; i < l1; ++i) { const int32_t n = sn[slist[i]]; cols += n > 0 ? (n + T - 1) / T : 1; }
                if (cols <= PA_NB) hi = T; else lo = T + 1;
            }
            sT[b] = lo;
        }
        __syncwarp();
        if (lane == 0) {
            long long steps = 0;
            int Tlong = 0;
            for (int b = 0; b < nb; ++b) { blockT[b] = sT[b]; step0[b] = steps; steps += sT[b]; Tlong = sT[b] > Tlong ? sT[b] : Tlong; }
            step0[nb] = steps;
            plan_out[0] = steps;
            plan_out[1] = nb;
            plan_out[2] = Tlong;
        }
    }
    __syncthreads();
    for (int i = tid; i < nb * PA_NB; i += blockDim.x) { col_group[i] = -1; col_meta[i] = -1; }
    __syncthreads();
    // columns of a block: its groups, c = ceil(n / T_b) each (one thread per block)
    if (tid < nb) {
        const int b = tid, T = sT[b];
        int fill = 0;
        for (int i = soff[b]; i < soff[b + 1]; ++i) {
            const int g = slist[i];
            const int32_t n = sn[g];
            const int c = (n > 0 && T > 0) ? (n + T - 1) / T : 1;
            for (int j = 0; j < c; ++j) { col_group[b * PA_NB + fill + j] = g; col_meta[b * PA_NB + fill + j] = j == c - 1 ? g : -2; }
            PaGroup r;
            r.goff0 = goff[g];
            const int64_t base = step0[b] * PA_NB + fill;
            r.base = (int32_t)(base > 0x7fffffff ? 0x7fffffff : base);
            r.c = c;
            grp[g] = r;
            col_last[g] = b * PA_NB + fill + c - 1;
            fill += c;
        }
    }
}

// ---- K1, scatter form: canonical normalise of the label-sorted raw segments straight into the interleaved layout ----
// Source ordered: a warp owns R consecutive raw rows, all of their 128-bit loads are issued before anything else, the
// (label -> group record) lookups of lanes 0..R-1 fly alongside; each row is written as one contiguous 2*Dp-byte run at
// its interleaved position.  Arithmetic = k_normalize_vec (canonical).
template <int NQ, typename TIn>
__global__ void __launch_bounds__(256)
k_pa_normalize_scatter(const TIn* __restrict__ x, const int32_t* __restrict__ lab, int32_t label_base, int64_t n, int32_t D, int32_t Dp,
                       const PaGroup* __restrict__ grp, __nv_bfloat16* __restrict__ out) {
    constexpr int R = (NQ <= 2) ? 4 : (NQ <= 4 ? 2 : 1);       // rows per warp iteration (register budget)
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
        int64_t my_dst = -1;
        if (lane < R && row0 + lane < n) {
            const int g = lab[row0 + lane] - label_base;
            const PaGroup pg = grp[g];
            const int32_t t = (int32_t)(row0 + lane - pg.goff0);
            my_dst = pg.c == 1 ? (int64_t)pg.base + (int64_t)t * PA_NB : (int64_t)pg.base + (int64_t)(t / pg.c) * PA_NB + (t % pg.c);
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
            const int64_t dst = __shfl_sync(0xffffffffu, my_dst, r);
            if (dst < 0) continue;                              // past the last row (uniform)
            uint2* orow = reinterpret_cast<uint2*>(out + dst * (int64_t)Dp);
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                int q = lane + 32 * i;
                if (q < nqp) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(__fmul_rn(v[r][i].x, inv), __fmul_rn(v[r][i].y, inv));
                    __nv_bfloat162 hi = __floats2bfloat162_rn(__fmul_rn(v[r][i].z, inv), __fmul_rn(v[r][i].w, inv));
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    orow[q] = pk;
                }
            }
            for (int q = lane + 32 * NQ; q < nqp; q += 32) orow[q] = make_uint2(0u, 0u);
        }
    }
}
// zero rows of the layout: steps of a column beyond its last segment (and every step of an unused column); warp per column
__global__ void __launch_bounds__(256)
k_pa_zero_pad(const int32_t* __restrict__ col_group, const int32_t* __restrict__ blockT, const int64_t* __restrict__ step0,
              const int64_t* __restrict__ goff, const PaGroup* __restrict__ grp, int32_t n_cols, int32_t Dp, __nv_bfloat16* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int col = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (col >= n_cols) return;
    const int b = col / PA_NB, slot = col % PA_NB;
    const int32_t Tb = blockT[b];
    const int64_t s0 = step0[b];
    const int g = col_group[col];
    int32_t len = 0;
    if (g >= 0) {
        const PaGroup pg = grp[g];
        const int64_t n = goff[g + 1] - goff[g];
        const int j = slot - (int)((int64_t)pg.base - s0 * PA_NB);          // part index of this column
        len = n > j ? (int32_t)((n - j + pg.c - 1) / pg.c) : 0;
    }
    const int n16 = Dp >> 3;                                                // 16-byte pieces per row
    for (int32_t t = len; t < Tb; ++t) {
        uint4* orow = reinterpret_cast<uint4*>(out + ((s0 + t) * PA_NB + slot) * (int64_t)Dp);
        for (int q = lane; q < n16; q += 32) orow[q] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------------
struct PaParams {
    PgParams pg;                   // slot arrays, P, tau, RB, mode, g_base (= first COLUMN of this batch)
    const int32_t* col_meta;       // [n_blocks*256] per column: label group g when it is the LAST column of g, -2 = inner
                                   //                column of a group (more follow), -1 = unused column
    const int32_t* blockT;         // [n_blocks] steps of each block
    const int64_t* step0;          // [n_blocks+1] first step of each block
    int32_t block_lo, n_blocks;    // blocks [block_lo, block_lo + n_blocks) of this batch
    int32_t sched_groups;          // k_poolacc2: 1 = static GROUP schedule (PaSched), 0 = units dealt round robin
};

// ---- static unit schedule of the CTA pairs (k_poolacc2) ----------------------------------------------------------------
// The RB units of a block stream the same slabs of the interleaved segment matrix, so they should run at the same time
// on different pairs: L2 then serves each slab RB times for one DRAM read.  Dealing the units round robin over the pairs
// (u -> pair u % npairs) only does that while npairs is a multiple of RB: with 74 pairs and 20 units per block the set of
// pairs that share a block changes every round, pairs arrive from blocks of different lengths, drift apart, and the slabs
// are fetched from DRAM three times (ncu, config 3: 23.6 GB per launch; 9.7 GB with 60 pairs, 8.0 GB with 40).
// GROUP schedule: pairs [g*RB, (g+1)*RB) form group g and always work on ONE block together (pair i of the group owns row
// block i).  The rem = npairs % RB left-over pairs form a last group; with d = gcd(RB, rem) it finishes Bl = rem / d blocks in
// R = RB / d rounds when their Bl * RB units are dealt round robin over its rem pairs (every pair busy in every round), while
// every full group finishes R blocks.  A super-round is therefore S = n_full * R + Bl blocks in R rounds with no idle pair
// (config 3: 3 groups of 20 + one of 14: R = 10, Bl = 7, S = 37).
struct PaSched {
    int32_t RB, n_blocks, npairs, n_full, rem, R, Bl, S, groups;
    int32_t p;                     // this pair
    int64_t i;                     // full groups / round robin: units taken so far; last group: super-round
    int32_t t;                     // last group: units taken in this super-round
    __device__ __forceinline__ void init(int32_t RB_, int32_t n_blocks_, int32_t npairs_, int32_t pair, int32_t groups_) {
        RB = RB_; n_blocks = n_blocks_; npairs = npairs_; p = pair; i = 0; t = 0;
        n_full = npairs / RB;
        rem = npairs - n_full * RB;
        int32_t d = RB, e = rem;
        while (e) { const int32_t r = d % e; d = e; e = r; }       // gcd(RB, rem); rem == 0 -> d = RB
        R = RB / d;
        Bl = rem / d;
        S = n_full * R + Bl;
        groups = (groups_ && n_full >= 1) ? 1 : 0;
    }
    __device__ __forceinline__ bool next(int32_t& bl, int32_t& rb) {
        if (!groups) {                                               // round robin over all units, block major
            const int64_t u = (int64_t)p + i * npairs;
            if (u >= (int64_t)n_blocks * RB) return false;
            bl = (int32_t)(u / RB);
            rb = (int32_t)(u - (int64_t)bl * RB);
            ++i;
            return true;
        }
        if (p < n_full * RB) {                                       // a full group: one unit of every block the group takes
            const int32_t g = p / RB;
            const int64_t sr = i / R, r = i - sr * R;
            const int64_t b = sr * S + r * n_full + g;
            if (b >= n_blocks) return false;
            bl = (int32_t)b;
            rb = p - g * RB;
            ++i;
            return true;
        }
        const int32_t q = p - n_full * RB;                          // the last group: its Bl blocks dealt round robin over rem pairs
        for (;;) {
            const int64_t base = i * S + (int64_t)n_full * R;
            if (base >= n_blocks) return false;
            const int32_t v = q + t * rem;
            if (v < Bl * RB) {
                const int64_t b = base + v / RB;
                ++t;
                if (b >= n_blocks) continue;                        // (a partial last super-round)
                bl = (int32_t)b;
                rb = v % RB;
                return true;
            }
            ++i;
            t = 0;
        }
    }
};

// KH = passes over K: with KH == 2 only half of a unit's bank tiles (KCL = ceil(KCH / 2) K chunks of both row tiles) are
// resident at a time -- pass 0 runs all T steps over the first K half, pass 1 over the second, into the same accumulators --
// so that D > 256 still gets TWO row tiles per streamed B stage (half the L2 -> SM operand stream of MT == 1).
template <int KCH, int MT, int STAGES, int KH>
__global__ void __launch_bounds__(MT == 2 ? PG_THREADS2 : PG_THREADS, 1)
k_poolacc(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ PaParams q) {
    constexpr int NC = PA_NB;
    constexpr int KCL = (KCH + KH - 1) / KH;                // K chunks resident at a time
    constexpr uint32_t A_TILE = 128 * 128;
    constexpr uint32_t A_BYTES = MT * KCL * A_TILE;
    constexpr uint32_t B_STAGE = NC * 128;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t NSLOT = 512 / (MT * NC);            // accumulator sets in TMEM: 2 when one row tile per unit (MT == 1)
    static_assert(MT * NC <= 512, "TMEM columns");
    const PgParams& p = q.pg;

    extern __shared__ uint8_t pg_smem_raw[];
    const uint32_t raw = pg_smem_u32(pg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + A_BYTES;
    const uint32_t sBar = sB + STAGES * B_STAGE;
    // bank tiles are handed over per K chunk: chunk kc of the NEXT unit is loaded as soon as the last step of this unit
    // has consumed chunk kc, so only the tail of the reload is exposed
    const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8 * KCL;
    const uint32_t bar_b_full = bar_a_empty + 8 * KCL, bar_b_empty = bar_b_full + 8 * STAGES;
    const uint32_t bar_t_full = bar_b_empty + 8 * STAGES, bar_t_empty = bar_t_full + 8 * NSLOT;
    const uint32_t s_tmem = bar_t_empty + 8 * NSLOT;
    uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(pg_smem_raw + (s_tmem - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int kc = 0; kc < KCL; ++kc) { pg_mbar_init(bar_a_full + 8 * kc, 1); pg_mbar_init(bar_a_empty + 8 * kc, 1); }
        for (int s = 0; s < STAGES; ++s) { pg_mbar_init(bar_b_full + 8 * s, 1); pg_mbar_init(bar_b_empty + 8 * s, 1); }
        for (uint32_t i = 0; i < NSLOT; ++i) { pg_mbar_init(bar_t_full + 8 * i, 1); pg_mbar_init(bar_t_empty + 8 * i, 4 * MT); }
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
            uint32_t stage = 0, phase = 0, a_bits = 0;              // a_bits: phase of every bank-tile barrier (bit per chunk slot)
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int32_t bl = (int32_t)(u / p.RB), rb = (int32_t)(u - (int64_t)bl * p.RB);
                const int32_t b = q.block_lo + bl;
                const int32_t T = q.blockT[b];
                if (T <= 0) continue;
                const int64_t s0 = q.step0[b];
#pragma unroll 1
                for (int h = 0; h < KH; ++h) {
                    const int kc0 = h * KCL, kc1 = kc0 + KCL < KCH ? kc0 + KCL : KCH;
                    for (int32_t t = 0; t < T; ++t) {
                        const int32_t crow = (int32_t)((s0 + t) * NC);
#pragma unroll 1
                        for (int kc = kc0; kc < kc1; ++kc) {
                            const int kcl = kc - kc0;
                            if (t == 0) {   // this pass's bank tiles, chunk by chunk, in the order the MMAs will want them
                                pg_mbar_wait(bar_a_empty + 8 * kcl, ((a_bits >> kcl) & 1u) ^ 1u);
                                pg_mbar_expect_tx(bar_a_full + 8 * kcl, MT * A_TILE);
#pragma unroll 1
                                for (int rt = 0; rt < MT; ++rt)
                                    pg_tma_load_2d(sA + (rt * KCL + kcl) * A_TILE, &tmapA, kc * 64,
                                                   (int32_t)((int64_t)rb * MT * 128 + rt * 128), bar_a_full + 8 * kcl);
                                a_bits ^= 1u << kcl;
                            }
                            pg_mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
                            pg_mbar_expect_tx(bar_b_full + 8 * stage, B_STAGE);
                            pg_tma_load_2d(sB + stage * B_STAGE, &tmapB, kc * 64, crow, bar_b_full + 8 * stage);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: all steps of the block accumulate into the same MT tiles =================
        // (the whole warp runs the loop, one elected lane issues: see pg_elect_one)
        {
            uint32_t stage = 0, phase = 0, a_bits = 0, uidx = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int32_t bl = (int32_t)(u / p.RB);
                const int32_t T = q.blockT[q.block_lo + bl];
                if (T <= 0) continue;
                const uint32_t slot = uidx % NSLOT, use = uidx / NSLOT;
                pg_mbar_wait(bar_t_empty + 8 * slot, (use & 1u) ^ 1u);      // this accumulator set has been read out
                pg_fence_after();
                const uint32_t td = tmem_base + slot * (MT * NC);
#pragma unroll 1
                for (int h = 0; h < KH; ++h) {
                    const int kc0 = h * KCL, kc1 = kc0 + KCL < KCH ? kc0 + KCL : KCH;
                    for (int32_t t = 0; t < T; ++t) {
#pragma unroll 1
                        for (int kc = kc0; kc < kc1; ++kc) {
                            const int kcl = kc - kc0;
                            if (t == 0) {
                                pg_mbar_wait(bar_a_full + 8 * kcl, (a_bits >> kcl) & 1u);
                                a_bits ^= 1u << kcl;
                            }
                            pg_mbar_wait(bar_b_full + 8 * stage, phase);
                            pg_fence_after();
                            const uint64_t db = pg_make_desc(sB + stage * B_STAGE);
                            if (pg_elect_one()) {
#pragma unroll
                                for (int rt = 0; rt < MT; ++rt) {
                                    const uint64_t da = pg_make_desc(sA + (rt * KCL + kcl) * A_TILE);
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        pg_mma_bf16(td + rt * NC, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC,
                                                    (h | t | kc | kk) != 0 ? 1u : 0u);
                                }
                                pg_commit(bar_b_empty + 8 * stage);
                                if (t == T - 1) pg_commit(bar_a_empty + 8 * kcl);   // the next pass's / unit's chunk may be loaded
                            }
                            __syncwarp();
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
                if (pg_elect_one()) pg_commit(bar_t_full + 8 * slot);
                __syncwarp();
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
            // lane l holds the column records of columns l, l+32, ..: 8 coalesced loads instead of 256 serial ones
            int32_t gid[NC / 32];
            float ginv[NC / 32];
#pragma unroll
            for (int i = 0; i < NC / 32; ++i) {
                const int g = q.col_meta[b * NC + i * 32 + lane];
                gid[i] = g;
                const int64_t n = g >= 0 ? p.goff[g + 1] - p.goff[g] : 0;
                ginv[i] = n > 0 ? 1.0f / (float)n : 0.f;
            }
            const uint32_t slot = uidx % NSLOT, use = uidx / NSLOT;
            pg_mbar_wait(bar_t_full + 8 * slot, use & 1u);
            pg_fence_after();
            const uint32_t td = tmem_base + slot * (MT * NC);
            float run = 0.f;                                                          // sum over the columns of one group
#pragma unroll
            for (int blk = 0; blk < NC / 32; ++blk) {
                float v[32];
                pg_tmem_ld32(td + lane_base + rt * NC + blk * 32, v);
                pg_tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    const int g = __shfl_sync(0xffffffffu, gid[blk], cc);
                    const float inv = __shfl_sync(0xffffffffu, ginv[blk], cc);
                    if (g == -1) continue;                                            // unused column (uniform)
                    run += v[cc];
                    if (g < 0) continue;                                              // inner column: the group goes on
                    const float val = run * inv;
                    run = 0.f;
                    if (inv == 0.f) continue;                                         // empty group
                    if (p.mode == 1) {
                        if (row < p.P) p.dense_out[row * (int64_t)p.dense_ld + g] = val;
                        continue;
                    }
                    bool pass = (val >= p.tau) && (row < p.P);
                    const int64_t pos = (int64_t)b * NC + blk * 32 + cc - p.g_base;   // column inside the batch
                    if (p.kth) {                                                      // running k-th best (low thresholds only)
                        const uint32_t thr = pg_kth_load(p, pos, lane);
                        const uint32_t key = sdk_fkey(val);
                        pg_kth_update(p, pos, tile128 * 4 + wq, pass ? key : 0u, thr, lane);
                        pass = pass && key >= thr;
                    }
                    const uint32_t mpass = __ballot_sync(0xffffffffu, pass);
                    if (mpass == 0) continue;                                         // slot counts were zeroed before the launch
                    pg_flush_write_call(&p, val, pass, mpass, pos * nsub + tile128 * 4 + wq, lane, row);
                }
            }
            pg_fence_before();
            __syncwarp();
            if (lane == 0) pg_mbar_arrive(bar_t_empty + 8 * slot);
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

// =================================================================================================
// 2-CTA variant (cta_group::2, option "cta_group" = 2): a cluster of two CTAs issues 256 x 256 x 16 MMAs.  CTA r owns bank
// rows [.., r*128 .. r*128+127] of every 256-row tile (its half of A and its own accumulators) and HALF of every streamed
// slab of the interleaved segment matrix (B is split along N across the pair): per CTA the L2 -> SM traffic of the streamed
// operand and its shared-memory footprint are halved (a B stage is 16 KB: the ring is twice as deep).  The leader CTA
// issues all MMAs; "full" barriers live in the leader and collect the TMA bytes of both CTAs; "empty" and accumulator-full
// barriers are signalled in both CTAs by multicast tcgen05.commit; the accumulator-empty barrier lives in the leader and
// is armed by the epilogue warps of both CTAs.  Unit = (block of 256 columns, pair of row blocks = MT * 256 bank rows).
// =================================================================================================
template <int KCH, int MT, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MT == 2 ? PG_THREADS2 : PG_THREADS, 1)
k_poolacc2(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ PaParams q) {
    constexpr int NC = PA_NB;
    constexpr uint32_t A_TILE = 128 * 128;                  // this CTA's 128 rows x 64 bf16 of a 256-row tile
    constexpr uint32_t A_BYTES = MT * KCH * A_TILE;
    constexpr uint32_t B_HALF = (NC / 2) * 128;             // this CTA's half of a (step, K chunk) slab
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    constexpr uint32_t NSLOT = 512 / (MT * NC);
    static_assert(MT * NC <= 512, "TMEM columns");
    const PgParams& p = q.pg;

    extern __shared__ uint8_t pg_smem_raw[];
    const uint32_t raw = pg_smem_u32(pg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + A_BYTES;
    const uint32_t sBar = sB + STAGES * B_HALF;
    const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8 * KCH;
    const uint32_t bar_b_full = bar_a_empty + 8 * KCH, bar_b_empty = bar_b_full + 8 * STAGES;
    const uint32_t bar_t_full = bar_b_empty + 8 * STAGES, bar_t_empty = bar_t_full + 8 * NSLOT;
    const uint32_t s_tmem = bar_t_empty + 8 * NSLOT;
    uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(pg_smem_raw + (s_tmem - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = pg_cluster_rank();
    const bool leader = crank == 0;
    if (warp == 0 && lane == 0) {
        for (int kc = 0; kc < KCH; ++kc) { pg_mbar_init(bar_a_full + 8 * kc, 1); pg_mbar_init(bar_a_empty + 8 * kc, 1); }
        for (int s = 0; s < STAGES; ++s) { pg_mbar_init(bar_b_full + 8 * s, 1); pg_mbar_init(bar_b_empty + 8 * s, 1); }
        for (uint32_t i = 0; i < NSLOT; ++i) { pg_mbar_init(bar_t_full + 8 * i, 1); pg_mbar_init(bar_t_empty + 8 * i, 2 * 4 * MT); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    pg_fence_before();
    __syncthreads();
    pg_cluster_sync();                                      // peer barriers are initialised before anyone signals them
    pg_fence_after();
    const uint32_t tmem_base = *s_tmem_ptr;

    // (p.RB = pairs of row blocks, MT * 256 bank rows each; the units of this pair come from PaSched)
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t l_a_full = pg_mapa(bar_a_full, 0), l_b_full = pg_mapa(bar_b_full, 0), l_t_empty = pg_mapa(bar_t_empty, 0);

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_bits = 0;
            PaSched sch;
            sch.init(p.RB, q.n_blocks, (int32_t)npairs, (int32_t)pair, q.sched_groups);
            int32_t bl, rb;
            while (sch.next(bl, rb)) {
                const int32_t b = q.block_lo + bl;
                const int32_t T = q.blockT[b];
                if (T <= 0) continue;
                const int64_t s0 = q.step0[b];
                for (int32_t t = 0; t < T; ++t) {
                    const int32_t crow = (int32_t)((s0 + t) * NC + crank * (NC / 2));
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc) {
                        if (t == 0) {   // this unit's bank tiles, chunk by chunk, in the order the MMAs will want them
                            pg_mbar_wait(bar_a_empty + 8 * kc, ((a_bits >> kc) & 1u) ^ 1u);
                            if (leader) pg_mbar_expect_tx(bar_a_full + 8 * kc, 2 * MT * A_TILE);
#pragma unroll 1
                            for (int rt = 0; rt < MT; ++rt)
                                pg_tma_load_2d_2sm(sA + (rt * KCH + kc) * A_TILE, &tmapA, kc * 64,
                                                   (int32_t)(((int64_t)rb * MT + rt) * 256 + crank * 128), l_a_full + 8 * kc);
                            a_bits ^= 1u << kc;
                        }
                        pg_mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
                        if (leader) pg_mbar_expect_tx(bar_b_full + 8 * stage, 2 * B_HALF);
                        pg_tma_load_2d_2sm(sB + stage * B_HALF, &tmapB, kc * 64, crow, l_b_full + 8 * stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only; the whole warp runs the loop, one elected lane issues) =================
        if (leader) {
            uint32_t stage = 0, phase = 0, a_bits = 0, uidx = 0;
            PaSched sch;
            sch.init(p.RB, q.n_blocks, (int32_t)npairs, (int32_t)pair, q.sched_groups);
            int32_t bl, rb;
            while (sch.next(bl, rb)) {
                const int32_t T = q.blockT[q.block_lo + bl];
                if (T <= 0) continue;
                const uint32_t slot = uidx % NSLOT, use = uidx / NSLOT;
                pg_mbar_wait(bar_t_empty + 8 * slot, (use & 1u) ^ 1u);      // both CTAs have read this accumulator set out
                pg_fence_after();
                const uint32_t td = tmem_base + slot * (MT * NC);
                for (int32_t t = 0; t < T; ++t) {
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc) {
                        if (t == 0) {
                            pg_mbar_wait(bar_a_full + 8 * kc, (a_bits >> kc) & 1u);
                            a_bits ^= 1u << kc;
                        }
                        pg_mbar_wait(bar_b_full + 8 * stage, phase);
                        pg_fence_after();
                        const uint64_t db = pg_make_desc(sB + stage * B_HALF);
                        if (pg_elect_one()) {
#pragma unroll
                            for (int rt = 0; rt < MT; ++rt) {
                                const uint64_t da = pg_make_desc(sA + (rt * KCH + kc) * A_TILE);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    pg_mma_bf16_2sm(td + rt * NC, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC,
                                                    (t | kc | kk) != 0 ? 1u : 0u);
                            }
                            pg_commit_2sm(bar_b_empty + 8 * stage);
                            if (t == T - 1) pg_commit_2sm(bar_a_empty + 8 * kc);   // the next unit's chunk may be loaded (both CTAs)
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                if (pg_elect_one()) pg_commit_2sm(bar_t_full + 8 * slot);
                __syncwarp();
                ++uidx;
            }
        }
    } else {
        // ================= epilogue (both CTAs): warpgroup wg <-> row tile; thread <-> one of this CTA's 128 bank rows =========
        const int wq = warp & 3;
        const int wg = (warp - 2) >> 2;
        const int rt = MT == 2 ? wg : 0;
        const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
        uint32_t uidx = 0;
        PaSched sch;
        sch.init(p.RB, q.n_blocks, (int32_t)npairs, (int32_t)pair, q.sched_groups);
        int32_t bl, rb;
        while (sch.next(bl, rb)) {
            const int32_t b = q.block_lo + bl;
            if (q.blockT[b] <= 0) continue;
            const int64_t tile128 = ((int64_t)rb * MT + rt) * 2 + crank;              // index of this CTA's 128-row tile
            const int64_t row = tile128 * 128 + wq * 32 + lane;
            const int64_t nsub = (int64_t)p.RB * MT * 8;
            // lane l holds the column record of column blk * 32 + l; the records of the next 32 columns are fetched while this
            // block of 32 is processed.  The loop over the eight blocks is ROLLED: fully unrolled, the 256 copies of the
            // per-column code (60 KB of instructions) streamed through the instruction cache once per unit -- invisible behind
            // hundreds of accumulation steps (config 3), but 70 us per unit when a unit is two steps (pool-first centroids).
            auto col_record = [&](int blk, int32_t& g_out, float& inv_out) {
                const int g = q.col_meta[b * NC + blk * 32 + lane];
                const int64_t n = g >= 0 ? p.goff[g + 1] - p.goff[g] : 0;
                g_out = g;
                inv_out = n > 0 ? 1.0f / (float)n : 0.f;
            };
            int32_t gid_c, gid_n = -1;
            float ginv_c, ginv_n = 0.f;
            col_record(0, gid_c, ginv_c);
            const uint32_t slot = uidx % NSLOT, use = uidx / NSLOT;
            pg_mbar_wait(bar_t_full + 8 * slot, use & 1u);
            pg_fence_after();
            const uint32_t td = tmem_base + slot * (MT * NC);
            float run = 0.f;                                                          // sum over the columns of one group
#pragma unroll 1
            for (int blk = 0; blk < NC / 32; ++blk) {
                if (blk + 1 < NC / 32) col_record(blk + 1, gid_n, ginv_n);
                float v[32];
                pg_tmem_ld32(td + lane_base + rt * NC + blk * 32, v);
                pg_tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    const int g = __shfl_sync(0xffffffffu, gid_c, cc);
                    const float inv = __shfl_sync(0xffffffffu, ginv_c, cc);
                    if (g == -1) continue;                                            // unused column (uniform)
                    run += v[cc];
                    if (g < 0) continue;                                              // inner column: the group goes on
                    const float val = run * inv;
                    run = 0.f;
                    if (inv == 0.f) continue;                                         // empty group
                    if (p.mode == 1) {
                        if (row < p.P) p.dense_out[row * (int64_t)p.dense_ld + g] = val;
                        continue;
                    }
                    bool pass = (val >= p.tau) && (row < p.P);
                    const int64_t pos = (int64_t)b * NC + blk * 32 + cc - p.g_base;   // column inside the batch
                    if (p.kth) {                                                      // running k-th best (low thresholds only)
                        const uint32_t thr = pg_kth_load(p, pos, lane);
                        const uint32_t key = sdk_fkey(val);
                        pg_kth_update(p, pos, tile128 * 4 + wq, pass ? key : 0u, thr, lane);
                        pass = pass && key >= thr;
                    }
                    const uint32_t mpass = __ballot_sync(0xffffffffu, pass);
                    if (mpass == 0) continue;                                         // slot counts were zeroed before the launch
                    pg_flush_write_call(&p, val, pass, mpass, pos * nsub + tile128 * 4 + wq, lane, row);
                }
                gid_c = gid_n;
                ginv_c = ginv_n;
            }
            pg_fence_before();
            __syncwarp();
            if (lane == 0) pg_mbar_arrive_cluster(l_t_empty + 8 * slot);
            ++uidx;
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

// ---- host side ----------------------------------------------------------------------------------------------------
template <int KCH, int MT, int STAGES, int KH>
static int pa_launch_t(sdk_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const PaParams& q, int grid) {
    constexpr int KCL = (KCH + KH - 1) / KH;
    constexpr size_t smem = (size_t)MT * KCL * 16384 + (size_t)STAGES * PA_NB * 128 + 320 + 1024;
    static_assert(smem <= PG_SMEM_LIMIT, "shared memory budget");
    auto kern = k_poolacc<KCH, MT, STAGES, KH>;
    SDK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, MT == 2 ? PG_THREADS2 : PG_THREADS, smem, c->stream>>>(ta, tb, q);
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}

static int pa_mt_for(int kch) { return kch <= 4 ? 2 : 1; }

template <int KCH, int MT, int STAGES>
static int pa_launch2_t(sdk_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const PaParams& q, int grid) {
    constexpr size_t smem = (size_t)MT * KCH * 16384 + (size_t)STAGES * (PA_NB / 2) * 128 + 384 + 1024;
    static_assert(smem <= PG_SMEM_LIMIT, "shared memory budget");
    auto kern = k_poolacc2<KCH, MT, STAGES>;
    SDK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, MT == 2 ? PG_THREADS2 : PG_THREADS, smem, c->stream>>>(ta, tb, q);      // __cluster_dims__(2,1,1): grid must be even
    c->launches++;
    SDK_CUDA(c, cudaGetLastError());
    return SDK_OK;
}
// cta_group::2: B stages are half as large, so the ring is twice as deep in the same shared memory
static int pa_launch2(sdk_ctx* c, int kch, const CUtensorMap& ta, const CUtensorMap& tb, const PaParams& q, int grid) {
    switch (kch) {
        case 1: return pa_launch2_t<1, 2, 8>(c, ta, tb, q, grid);
        case 2: return pa_launch2_t<2, 2, 8>(c, ta, tb, q, grid);
        case 3: return pa_launch2_t<3, 2, 8>(c, ta, tb, q, grid);
        case 4: return pa_launch2_t<4, 2, 6>(c, ta, tb, q, grid);
        case 5: return pa_launch2_t<5, 1, 8>(c, ta, tb, q, grid);
        case 6: return pa_launch2_t<6, 1, 8>(c, ta, tb, q, grid);
        case 7: return pa_launch2_t<7, 1, 6>(c, ta, tb, q, grid);
        default: return pa_launch2_t<8, 1, 6>(c, ta, tb, q, grid);
    }
}

// (KCH, MT, STAGES, KH).  Shared memory = MT * ceil(KCH / KH) * 16 KB (bank tiles) + STAGES * 32 KB (B ring).  Above 256
// dimensions one row tile per unit with all K chunks resident measured FASTER (config 4 forced: 10.8 ms) than two row
// tiles with the K halves resident in turn (<8, 2, 3, 2>: 12.3 ms, although it halves the L2 -> SM operand stream), so
// the KH = 2 instantiations are not used.
static int pa_launch(sdk_ctx* c, int kch, const CUtensorMap& ta, const CUtensorMap& tb, const PaParams& q, int grid) {
    switch (kch) {
        case 1: return pa_launch_t<1, 2, 6, 1>(c, ta, tb, q, grid);
        case 2: return pa_launch_t<2, 2, 5, 1>(c, ta, tb, q, grid);
        case 3: return pa_launch_t<3, 2, 4, 1>(c, ta, tb, q, grid);
        case 4: return pa_launch_t<4, 2, 3, 1>(c, ta, tb, q, grid);
        case 5: return pa_launch_t<5, 1, 4, 1>(c, ta, tb, q, grid);
        case 6: return pa_launch_t<6, 1, 4, 1>(c, ta, tb, q, grid);
        case 7: return pa_launch_t<7, 1, 3, 1>(c, ta, tb, q, grid);
        default: return pa_launch_t<8, 1, 3, 1>(c, ta, tb, q, grid);
    }
}

int sdk_poolacc_applicable(int32_t Dp, int32_t G, int32_t pool) {
    return sdk_poolgemm_supported(Dp) && pool == SDK_POOL_MEAN && G >= 1;
}

// Plan of the interleaved layout (see the header).  Returns the total number of 256-row steps in *steps_out (one
// stream sync); S*256 / N - 1 is the zero padding.  The plan stays in the context until the next call.
int sdk_poolacc_plan(sdk_ctx* c, const int64_t* d_goff, int32_t G, int64_t N, int64_t P, int32_t Dp, int64_t* steps_out) {
    const bool small = G <= PA_SMALL_G;
    const int kch = Dp / 64, MT = pa_mt_for(kch);
    int32_t t_min = 8;
    if (N / ((int64_t)PA_NB * 2048) > t_min) t_min = (int32_t)(N / ((int64_t)PA_NB * 2048));
    int64_t max_blocks64 = small ? 2 * (((int64_t)G + N / t_min) / PA_NB + 1) + 2 : ((int64_t)G + PA_NB - 1) / PA_NB;
    if (small && max_blocks64 < PA_MAX_BINS) max_blocks64 = PA_MAX_BINS;
    if (max_blocks64 > (1 << 22)) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling plan: too many column blocks");
    const int32_t max_blocks = (int32_t)max_blocks64;
    SDK_TRY(sdk_reserve(c, c->pa_col_group, (size_t)max_blocks * PA_NB * 4));
    SDK_TRY(sdk_reserve(c, c->pa_blockT, (size_t)max_blocks * 4));
    SDK_TRY(sdk_reserve(c, c->pa_step0, (size_t)(max_blocks + 4) * 8));
    SDK_TRY(sdk_reserve(c, c->pa_grp, (size_t)G * sizeof(PaGroup)));
    SDK_TRY(sdk_reserve(c, c->pa_col_last, (size_t)G * 4));
    int64_t* step0 = (int64_t*)c->pa_step0.p;
    int64_t* plan_out = step0 + max_blocks + 1;                     // {steps, blocks}
    if (small) {
        SDK_TRY(sdk_reserve(c, c->pa_col_meta, (size_t)max_blocks * PA_NB * 4));
        PaCost pc;
        pc.step_cycles = (float)(kch * 4 * MT * 128);
        pc.unit_cycles = 6000.f + (float)(MT * kch) * 1400.f;       // TMEM read-out + bank tile load (16 KB tiles over TMA)
        pc.RB = (int32_t)((P + (int64_t)MT * 128 - 1) / ((int64_t)MT * 128));
        pc.sms = c->sm_count;
        pc.t_min = t_min;
        sdk_prof_scope ps(c, "plan");
        if (G <= PA_BINS_G)
            k_pa_plan_bins<<<1, 1024, 0, c->stream>>>(d_goff, G, pc, (int32_t*)c->pa_col_group.p, (int32_t*)c->pa_col_meta.p,
                                                      (int32_t*)c->pa_blockT.p, step0, plan_out, (PaGroup*)c->pa_grp.p,
                                                      (int32_t*)c->pa_col_last.p);
        else
            k_pa_plan_small<<<1, 1024, 0, c->stream>>>(d_goff, G, pc, max_blocks, (int32_t*)c->pa_col_group.p, (int32_t*)c->pa_col_meta.p,
                                                       (int32_t*)c->pa_blockT.p, step0, plan_out, (PaGroup*)c->pa_grp.p,
                                                       (int32_t*)c->pa_col_last.p);
        c->launches += 1;
        SDK_CUDA(c, cudaGetLastError());
    } else {
        SDK_TRY(sdk_reserve(c, c->pa_hist, (size_t)PA_BUCKETS * 4));
        SDK_TRY(sdk_reserve(c, c->pa_sorted, (size_t)G * 4));
        SDK_TRY(sdk_reserve(c, c->pa_pos, (size_t)G * 4));
        int32_t* hist = (int32_t*)c->pa_hist.p;
        sdk_prof_scope ps(c, "plan");
        SDK_CUDA(c, cudaMemsetAsync(hist, 0, (size_t)PA_BUCKETS * 4, c->stream));
        k_pa_hist<<<(G + 255) / 256, 256, 0, c->stream>>>(d_goff, G, hist);
        k_pa_scan<<<1, 1024, 0, c->stream>>>(hist);
        k_pa_scatter<<<(G + 255) / 256, 256, 0, c->stream>>>(d_goff, G, hist, (int32_t*)c->pa_sorted.p, (int32_t*)c->pa_pos.p);
        k_pa_blocks<<<(max_blocks + 127) / 128, 128, 0, c->stream>>>(d_goff, (const int32_t*)c->pa_sorted.p, G, max_blocks,
                                                                     (int32_t*)c->pa_col_group.p, (int32_t*)c->pa_blockT.p);
        k_pa_steps<<<1, 32, 0, c->stream>>>((const int32_t*)c->pa_blockT.p, max_blocks, step0, plan_out);
        k_pa_group_rows<<<(G + 255) / 256, 256, 0, c->stream>>>(d_goff, (const int32_t*)c->pa_pos.p, step0, G, (PaGroup*)c->pa_grp.p,
                                                                (int32_t*)c->pa_col_last.p);
        c->launches += 6;
        SDK_CUDA(c, cudaGetLastError());
    }
    int64_t h[3] = {0, 0, 0};
    SDK_CUDA(c, cudaMemcpyAsync(h, plan_out, 24, cudaMemcpyDeviceToHost, c->stream));
    SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h[1] < 0 || h[1] > max_blocks) return sdk_fail(c, SDK_ECUDA, "accumulate-pooling plan: inconsistent block count");
    c->pa_blocks = (int32_t)h[1];
    c->pa_split = small;
    c->pa_chain_max = h[2];
    *steps_out = h[0];
    return SDK_OK;
}

// The GEMM (+ merge) over a prepared interleaved matrix and the plan in c->pa_* (n_blocks = c->pa_blocks), in batches of
// blocks that keep the candidate slots under ~6 GB.
static int pa_gemm(sdk_ctx* c, const void* il_p, int64_t n_rows, int32_t Dp, const __nv_bfloat16* d_rows, int64_t P, const int64_t* d_goff,
                   int32_t G, int32_t mode, float tau, int32_t ncand, int32_t* d_cand_row, float* d_gbound, float* d_dense) {
    const int kch = Dp / 64, MT = pa_mt_for(kch);
    const int32_t n_blocks = c->pa_blocks;
    const int64_t* step0 = (const int64_t*)c->pa_step0.p;
    const int32_t* col_group = (const int32_t*)c->pa_col_group.p;
    const int32_t* col_meta = c->pa_split ? (const int32_t*)c->pa_col_meta.p : col_group;    // c == 1: every column is a last column
    // ---- GEMM (+ merge), in batches of blocks that keep the candidate slots under ~6 GB ----
    // option "cta_group" = 2: CTA pairs (k_poolacc2); a unit then covers MT * 256 bank rows and RB counts row-block PAIRS
    // (auto: many one-column groups -- config 3, the pool-first centroids; the column-split plans of a few large groups
    //  were tuned against the single-CTA schedule and stay there)
    const bool cta2 = c->sm_count >= 2 && (c->opt_cta_group == 2 || (c->opt_cta_group == 0 && mode == 0 && !c->pa_split && n_blocks >= 64));
    const int64_t rows_per_block = (int64_t)MT * (cta2 ? 256 : 128);
    const int32_t RB = (int32_t)((P + rows_per_block - 1) / rows_per_block);
    const int32_t nsub = RB * MT * (cta2 ? 8 : 4);
    const size_t per_group = (size_t)nsub * (PG_CS * 8 + 8);
    int64_t bbatch = n_blocks;
    if (mode == 0) {
        bbatch = (int64_t)((size_t)(6144ull << 20) / (per_group * PA_NB));
        if (bbatch < 1) bbatch = 1;
        if (bbatch > n_blocks) bbatch = n_blocks;
        SDK_TRY(sdk_reserve(c, c->slot_cnt, (size_t)bbatch * PA_NB * nsub * 4));
        SDK_TRY(sdk_reserve(c, c->slot_bound, (size_t)bbatch * PA_NB * nsub * 4));
        SDK_TRY(sdk_reserve(c, c->slot_row, (size_t)bbatch * PA_NB * nsub * PG_CS * 4));
        SDK_TRY(sdk_reserve(c, c->slot_val, (size_t)bbatch * PA_NB * nsub * PG_CS * 4));
    }
    CUtensorMap ta, tb;
    SDK_TRY(pg_make_tmap(c, &ta, d_rows, P, Dp, 128));
    SDK_TRY(pg_make_tmap(c, &tb, il_p, n_rows, Dp, cta2 ? PA_NB / 2 : PA_NB));
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
        q.pg.mode = mode;
        q.pg.slot_cnt = (int32_t*)c->slot_cnt.p;
        q.pg.slot_row = (int32_t*)c->slot_row.p;
        q.pg.slot_val = (float*)c->slot_val.p;
        q.pg.slot_bound = (float*)c->slot_bound.p;
        q.pg.dense_out = d_dense;
        q.pg.dense_ld = G;
        q.pg.nsub = nsub;
        q.pg.kth = nullptr;
        if (mode == 0 && c->kth_on) {
            SDK_TRY(sdk_reserve(c, c->kth, (size_t)bbatch * PA_NB * PG_KTH * 4));
            SDK_CUDA(c, cudaMemsetAsync(c->kth.p, 0, (size_t)(bb - ba) * PA_NB * PG_KTH * 4, c->stream));
            q.pg.kth = (uint32_t*)c->kth.p;
        }
        q.col_meta = col_meta;
        q.blockT = (const int32_t*)c->pa_blockT.p;
        q.step0 = step0;
        q.block_lo = (int32_t)ba;
        q.n_blocks = (int32_t)(bb - ba);
        q.sched_groups = 0;
        const int64_t n_units = (bb - ba) * RB;
        int grid = cta2 ? 2 * (int)std::min<int64_t>(n_units, c->sm_count / 2) : (int)std::min<int64_t>(n_units, c->sm_count);
        if (const char* e = getenv("SDK_PA_GRID")) {            // experiment knob (profiles/r02_poolacc_exp_*): fewer CTAs than SMs
            const int g = atoi(e);
            if (g >= 2 && g < grid) grid = cta2 ? (g & ~1) : g;
        }
        if (cta2) {
            // group schedule (PaSched) when the pairs split into groups of RB that stay busy: n_full groups take w blocks
            // while the rem left-over pairs finish one
            const int npairs = grid / 2, n_full = npairs / RB, rem = npairs - n_full * RB;
            if (n_full >= 1) {
                int d = RB, e = rem;
                while (e) { const int r = d % e; d = e; e = r; }
                const int S = n_full * (RB / d) + rem / d;                 // blocks per super-round (PaSched)
                q.sched_groups = q.n_blocks >= 4 * S ? 1 : 0;              // (a few blocks only: the left-over pairs would idle)
            }
            if (const char* e = getenv("SDK_PA_SCHED")) q.sched_groups = atoi(e) != 0 ? 1 : 0;     // A/B runs
        }
        // slots of columns that are never flushed (inner / unused columns, empty groups) must read as empty
        if (mode == 0) SDK_CUDA(c, cudaMemsetAsync(c->slot_cnt.p, 0, (size_t)(bb - ba) * PA_NB * nsub * 4, c->stream));
        {
            sdk_prof_scope ps(c, "poolgemm");
            if (cta2) SDK_TRY(pa_launch2(c, kch, ta, tb, q, grid));
            else SDK_TRY(pa_launch(c, kch, ta, tb, q, grid));
        }
        if (mode == 0) {
            pg_launch_merge(c, d_goff, (int32_t)(ba * PA_NB), (int32_t)((bb - ba) * PA_NB), nsub, col_meta, tau, ncand, d_cand_row, d_gbound);
            SDK_CUDA(c, cudaGetLastError());
        }
    }
    if (mode == 0 && bbatch >= n_blocks) {     // one batch: every label's candidate slots are still in memory (second chance)
        c->slot_g0 = 0;
        c->slot_g1 = G;
        c->slot_nsub = nsub;
        c->slot_by_col = true;
    }
    return SDK_OK;
}

// Normalises the raw segments into the planned interleaved layout (buffer `il`), runs the accumulate-pooling GEMM and,
// in candidate mode, the slot merge.  `S` = steps from sdk_poolacc_plan (same goff, no other plan in between).
int sdk_launch_poolacc(sdk_ctx* c, const void* d_seg_raw, int32_t in_dtype, const int32_t* d_seg_label, int32_t label_base, int64_t N, int32_t D,
                       int32_t Dp, const __nv_bfloat16* d_rows, int64_t P, const int64_t* d_goff, int32_t G, int64_t S, int32_t mode,
                       float tau, int32_t ncand, int32_t* d_cand_row, float* d_gbound, float* d_dense, sdk_buf& il,
                       const PaGroup** d_grp_out) {
    if (!sdk_poolacc_applicable(Dp, G, SDK_POOL_MEAN) || !c->tmap_encode) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path unavailable");
    if (P > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "tcgen05 path: at most 2^31-1 bank rows");
    if (D % 4 != 0 || D > 2048) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path needs D % 4 == 0");
    if ((uintptr_t)d_seg_raw % (in_dtype == SDK_IN_F16 ? 8 : 16) != 0)
        return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path needs a 16-byte (fp16 rows: 8-byte) aligned segment matrix");
    const int kch = Dp / 64, MT = pa_mt_for(kch);
    const int32_t n_blocks = c->pa_blocks;
    int64_t* step0 = (int64_t*)c->pa_step0.p;
    const int32_t* col_group = (const int32_t*)c->pa_col_group.p;
    const int32_t* col_meta = c->pa_split ? (const int32_t*)c->pa_col_meta.p : col_group;    // c == 1: every column is a last column
    const PaGroup* grp = (const PaGroup*)c->pa_grp.p;
    const int64_t n_rows = S * PA_NB;
    if (n_rows > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path: interleaved matrix exceeds 2^31-1 rows");
    if (d_grp_out) *d_grp_out = grp;
    if (S == 0) {
        if (mode == 0) SDK_CUDA(c, cudaMemsetAsync(d_cand_row, 0xff, (size_t)G * ncand * 4, c->stream));
        return SDK_OK;
    }
    SDK_TRY(sdk_reserve(c, il, (size_t)n_rows * Dp * 2));
    // ---- K1, scatter form + zero rows ----
    {
        sdk_prof_scope ps(c, "normalize");
        int nq = (D / 4 + 31) / 32;
        const int R = nq <= 2 ? 4 : (nq <= 4 ? 2 : 1);
        int64_t blocks64 = ((N + R - 1) / R + 7) / 8;
        if (blocks64 < 1) blocks64 = 1;
        int blocks = (int)(blocks64 < (int64_t)c->sm_count * 8 ? blocks64 : (int64_t)c->sm_count * 8);
#define PA_NORM(NQ)                                                                                                                   \
    do {                                                                                                                              \
        if (in_dtype == SDK_IN_F16)                                                                                                   \
            k_pa_normalize_scatter<NQ, __half><<<blocks, 256, 0, c->stream>>>((const __half*)d_seg_raw, d_seg_label, label_base, N, D, Dp, grp, (__nv_bfloat16*)il.p); \
        else                                                                                                                          \
            k_pa_normalize_scatter<NQ, float><<<blocks, 256, 0, c->stream>>>((const float*)d_seg_raw, d_seg_label, label_base, N, D, Dp, grp, (__nv_bfloat16*)il.p);   \
    } while (0)
        if (nq <= 1) PA_NORM(1);
        else if (nq <= 2) PA_NORM(2);
        else if (nq <= 4) PA_NORM(4);
        else if (nq <= 8) PA_NORM(8);
        else PA_NORM(16);
#undef PA_NORM
        const int32_t n_cols = n_blocks * PA_NB;
        k_pa_zero_pad<<<(n_cols + 7) / 8, 256, 0, c->stream>>>(col_group, (const int32_t*)c->pa_blockT.p, step0, d_goff, grp, n_cols, Dp,
                                                               (__nv_bfloat16*)il.p);
        c->launches += 2;
        SDK_CUDA(c, cudaGetLastError());
    }
    return pa_gemm(c, il.p, n_rows, Dp, d_rows, P, d_goff, G, mode, tau, ncand, d_cand_row, d_gbound, d_dense);
}

// Stage A over a matrix that is ALREADY in the interleaved layout with its plan in c->pa_* (pool-first: normalize.cu builds
// both for the label centroids -- two rows per label, one column per label, two steps per block).
int sdk_launch_poolacc_prepared(sdk_ctx* c, const void* d_il, int64_t n_rows, int32_t Dp, const __nv_bfloat16* d_rows, int64_t P,
                                const int64_t* d_goff, int32_t G, float tau, int32_t ncand, int32_t* d_cand_row, float* d_gbound) {
    if (!sdk_poolacc_applicable(Dp, G, SDK_POOL_MEAN) || !c->tmap_encode) return sdk_fail(c, SDK_EINVAL, "accumulate-pooling path unavailable");
    if (P > 0x7fffffffLL || n_rows > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "tcgen05 path: at most 2^31-1 rows");
    return pa_gemm(c, d_il, n_rows, Dp, d_rows, P, d_goff, G, 0, tau, ncand, d_cand_row, d_gbound, nullptr);
}
