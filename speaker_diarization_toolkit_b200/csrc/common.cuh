// common.cuh -- context, error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <map>
#include <vector>

#include "../../include/sdk_b200.h"

#define SDK_Q30 1073741824.0
#define SDK_NSTAGE 3

struct sdk_buf {               // grow-only device buffer
    void* p = nullptr;
    size_t cap = 0;
};

// Accumulate-pooling layout (poolacc.cu): segment t of a label group lives at row  base + (t / c) * 256 + (t % c)  of the
// group-interleaved bf16 matrix (c = accumulator columns the group is dealt over, round robin).
struct PaGroup {
    int64_t goff0;             // first segment of the group in the caller's (label-sorted) order
    int32_t base;              // row of segment 0 in the interleaved matrix
    int32_t c;                 // columns (parts) of the group, >= 1
};

struct sdk_prof_entry {
    double ms = 0.0;
    int64_t launches = 0;
};

struct sdk_pending_ev {
    std::string name;
    cudaEvent_t a, b;
};

// Result record of one rank.  ONE device allocation: the head [flags | rows | scores | spk | counts | trust] is exactly
// what the row-sharded mode all-gathers (no packing copies), the tail holds the assignment outputs, and a fetch of a
// small record is one device->host copy.
#define SDK_FLAG_LABEL 0     // bit 0: label out of range, bit 1: labels not sorted
#define SDK_FLAG_FB 1        // groups whose top-k certificate failed (counter, consumed by the host)
#define SDK_FLAG_STATUS 2    // != 0: this rank failed before the collective (-error code)
#define SDK_FLAG_PEER 3      // after the merge: 1 + first rank whose record carries a label flag / status
#define SDK_NFLAGS 16
struct sdk_out_view {
    int32_t* flags = nullptr;
    int64_t* row = nullptr;
    float* score = nullptr;
    int32_t* spk = nullptr;
    int32_t* count = nullptr;
    uint8_t* trust = nullptr;
    int32_t* as_idx = nullptr;
    double* as_score = nullptr;
    int32_t* as_conf = nullptr;
    int32_t* as_cidx = nullptr;
    double* as_cscore = nullptr;
    size_t off_row = 0, off_score = 0, off_spk = 0, off_count = 0, off_trust = 0, gather_bytes = 0;
    size_t off_as_score = 0, off_as_cscore = 0, off_as_idx = 0, off_as_conf = 0, off_as_cidx = 0, bytes = 0;
};

// What the small-query path (gemv.cu) leaves open when its identify call returns: see sdk_settle (api.cu).
struct sdk_lazy {
    bool active = false;
    const void* seg_ops = nullptr;
    int32_t pool = 0, k = 0;
    double threshold = 0.0;
    int64_t* o_row = nullptr;
    float* o_score = nullptr;
    int32_t* o_count = nullptr;
    uint8_t* o_trust = nullptr;
    int32_t* o_spk = nullptr;
};

struct sdk_ctx {
    int device = 0, world = 1, rank = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    std::string err;
    int sm_count = 148;
    // options
    int opt_path = 0;          // 0 auto, 1 exact, 2 tensor
    double opt_eps = -1.0;     // <0: default by dtype
    int opt_profile = 0;
    int opt_cand = 16;         // re-scored candidates per label group (tensor path)
    int opt_cta_group = 0;     // tcgen05 kernels: 1 = single CTA, 2 = CTA pairs (cta_group::2), 0 = auto: pairs for accumulate-pooling
                               // (+3 % on config 3 at equal power: half the L2 -> SM operand stream), single CTA for k_poolgemm
    int opt_acc = 1;           // mean pooling with many label groups: pool inside the MMA accumulation (poolacc.cu)
    int opt_gemv = 1;          // <= 8 query segments: stream the bank once on the CUDA cores (gemv.cu) instead of tcgen05 tiles
    int opt_chunk_mb = 128;    // host-buffer identify: H2D/compute pipeline chunk size
    int opt_kth = 1;           // candidate flush prunes against a running per-label 64th-best bound: 0 off, 1 auto (low thresholds), 2 on
    int opt_poolfirst = 0;     // mean pooling on the tensor path: stage A contracts the label CENTROIDS (a different algorithm: HBM-bound)
    int opt_inject_fail = 0;   // test knob (multi-rank error handling): fail the next local identify pass
    // bank
    int64_t P = 0;
    int32_t D = 0, Dp = 0, dtype = 0;
    int32_t in_dtype = 0;      // storage type of the raw segment rows of the identify call in progress (SDK_IN_*)
    int64_t row_offset = 0;
    sdk_buf bank_f32, bank_bf16, row_speaker, row_trust;
    // segments / scratch
    sdk_buf seg_raw, seg_lab, seg_f32, seg_bf16, goff, qpool, dense, flags;
    sdk_buf cand_row, cand_val, cand_cnt, gbound, slot_cnt, slot_row, slot_val, slot_bound, range_g;
    sdk_buf fb_list, fb_rows, fb_list2, cand_row2, qpool2;
    sdk_buf cent_sum, cent_seg, goff2;   // pool-first stage A: fp32 centroid sums [G, Dp], hi/lo pseudo-segments [2G, Dp], their offsets
    sdk_buf kth;               // running k-th best buckets of the candidate flush (tcgen05.cuh)
    bool kth_on = false;       // set per identify call: the candidate threshold is low enough for noise rows to pass
    int32_t slot_g0 = 0, slot_g1 = 0, slot_nsub = 0;   // label groups whose candidate slots are live after stage A
    bool slot_by_col = false;  // slots are indexed by accumulator column (accumulate-pooling): group -> pa_col_last
    sdk_buf stage_seg[SDK_NSTAGE], stage_lab[SDK_NSTAGE];     // host-buffer identify: H2D staging ring
    sdk_buf pa_hist, pa_sorted, pa_pos, pa_col_group, pa_col_meta, pa_blockT, pa_step0, pa_grp, pa_col_last, seg_il;   // accumulate-pooling plan + layout
    int32_t pa_blocks = 0;     // blocks of 256 accumulator columns in the current plan
    int64_t pa_chain_max = 0;  // longest accumulation chain of the plan, in segments per column (max T_b)
    bool pa_split = false;     // current plan deals groups over several columns (col_meta != col_group)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[SDK_NSTAGE] = {}, ev_consumed[SDK_NSTAGE] = {};
    // results of the last identify
    int32_t L = 0, k = 0;
    int64_t N = 0;
    bool have_results = false, have_assign = false;
    sdk_lazy lazy;             // pending certificate / label-flag check of a small-query identify
    double assign_thr = 0.0;   // parameters of the last sdk_assign (re-run when a pending certificate changes the lists)
    int32_t assign_min_trust = 99;
    sdk_buf out_pack;          // the result record (sdk_out_view)
    sdk_out_view out;
    void* h_pack = nullptr;    // pinned host copy of a small record (one D2H per fetch)
    size_t h_pack_cap = 0;
    sdk_buf gather;            // NCCL all-gather receive buffer: world records
    bool bank_ok = false;      // a bank (possibly an empty shard of a row-sharded one) has been loaded
    void* nccl_comm = nullptr;
    int last_path = 0;
    float last_eps_base = 0.f, last_eps_chain = 0.f;   // certificate margin model of the last tensor-path call (select.cu)
    int32_t last_ncand = 0, last_cand_groups = 0;       // shape of cand_row / cand_val left by the last stage A
    int64_t last_fallback = 0;   // label groups re-done exhaustively (certificate failed twice)
    int64_t last_retry = 0;      // label groups whose certificate needed the second, wider candidate list
    int64_t launches = 0;
    std::map<std::string, sdk_prof_entry> prof;
    std::vector<sdk_pending_ev> pending;
    void* tmap_encode = nullptr;   // cuTensorMapEncodeTiled
};

int sdk_fail(sdk_ctx* c, int code, const std::string& msg);
int sdk_reserve(sdk_ctx* c, sdk_buf& b, size_t bytes);

#define SDK_CUDA(c, expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return sdk_fail((c), _e == cudaErrorMemoryAllocation ? SDK_ENOMEM : SDK_ECUDA,  \
                            std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    } while (0)

#define SDK_TRY(expr)             \
    do {                          \
        int _r = (expr);          \
        if (_r != SDK_OK) return _r; \
    } while (0)

// per-kernel timing scope (only active with option "profile")
struct sdk_prof_scope {
    sdk_ctx* c;
    const char* name;
    cudaEvent_t a = nullptr, b = nullptr;
    sdk_prof_scope(sdk_ctx* c_, const char* n) : c(c_), name(n) {
        if (c->opt_profile) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, c->stream);
        }
    }
    ~sdk_prof_scope() {
        if (c->opt_profile) {
            cudaEventRecord(b, c->stream);
            c->pending.push_back({name, a, b});
        }
    }
};

// ---- launchers implemented in the .cu files -------------------------------------------------
// K1: canonical L2 normalise (+ optional fp32 copy, + optional zero-padded bf16 copy)
int sdk_launch_normalize(sdk_ctx* c, const float* d_x, int64_t n, int32_t D, int32_t Dp,
                         float* d_f32 /*[n,D] or null*/, __nv_bfloat16* d_bf16 /*[n,Dp] or null*/);
// same, raw rows stored as fp32 (SDK_IN_F32) or IEEE fp16 (SDK_IN_F16: widened exactly to fp32, then the same arithmetic)
int sdk_launch_normalize_in(sdk_ctx* c, const void* d_x, int32_t in_dtype, int64_t n, int32_t D, int32_t Dp,
                            float* d_f32, __nv_bfloat16* d_bf16);
// pool-first stage A (normalize.cu): K1 that also accumulates the label centroids, then the hi/lo bf16 split
int sdk_poolfirst_applicable(const void* d_x, int32_t in_dtype, int32_t D, int32_t Dp);
int sdk_launch_normalize_centroid(sdk_ctx* c, const void* d_x, int32_t in_dtype, const int32_t* d_lab, int32_t label_base, int64_t n,
                                  int32_t D, int32_t Dp, float* d_f32, __nv_bfloat16* d_bf16, const int64_t* d_goff, int32_t G,
                                  int32_t round_bf16, int64_t* n_rows_out);
// group offsets from sorted labels; *d_flag != 0 when labels are unsorted / out of range
int sdk_launch_group_offsets(sdk_ctx* c, const int32_t* d_lab, int64_t N, int32_t L, int32_t label_base,
                             int64_t* d_goff, int32_t* d_flag);
// exact canonical Q30 pooling.  rows: dense (cand_row == null: slot == bank row, nslot == P) or
// sparse (cand_row[g*nslot + j] = bank row or -1).  glist (may be null) maps launch group -> group.
// Segment t of group g lives at row grp[g].base + (t / grp[g].c) * 256 + t % grp[g].c of d_seg_ops (grp == null: goff[g] + t).
int sdk_launch_exact(sdk_ctx* c, const void* d_seg_ops, const void* d_bank_ops, int32_t is_bf16,
                     int32_t D, int32_t pitch, const int64_t* d_goff, const int32_t* d_glist,
                     int32_t ngroups, const int32_t* d_cand_row, int64_t nslot, int32_t pool,
                     long long* d_qpool /*[ngroups,nslot]*/, const PaGroup* d_grp = nullptr,
                     int64_t n_seg = -1 /* >= 0: clamp group offsets to [0, n_seg] (labels not validated yet) */);
// select: speaker dedupe + threshold + ordered top-k (+ certificate on the sparse path)
int sdk_launch_select(sdk_ctx* c, const long long* d_qpool, const int64_t* d_goff,
                      const int32_t* d_glist, int32_t ngroups, const int32_t* d_cand_row,
                      int64_t nslot, int32_t pool, const int32_t* d_row_speaker,
                      const uint8_t* d_row_trust, double threshold, int32_t k, int64_t row_offset,
                      const float* d_gbound /*null on dense*/, float eps_base, float eps_chain, const PaGroup* d_grp, int32_t chain_div,
                      int32_t* d_fb_count,
                      int32_t* d_fb_list, int64_t* d_out_row, float* d_out_score,
                      int32_t* d_out_count, uint8_t* d_out_trust, int32_t* d_out_spk);
int sdk_launch_assign(sdk_ctx* c, const int64_t* d_row, const float* d_score,
                      const uint8_t* d_trust, const int32_t* d_count, int32_t L, int32_t k,
                      double thr, int32_t min_trust, int32_t* d_idx, double* d_ascore,
                      int32_t* d_conf, int32_t* d_cidx, double* d_cscore);
// merge after the all-gather: `world` result records (layout of `v`, `stride` bytes apart, starting at d_all) -> v
int sdk_launch_merge_topk(sdk_ctx* c, const void* d_all, size_t stride, const sdk_out_view& v, int32_t world, int32_t L, int32_t k);
// rows -1 / counts 0 for the groups [0, L) of the record (a rank without bank rows, or one that failed)
int sdk_launch_fill_empty(sdk_ctx* c, const sdk_out_view& v, int32_t L, int32_t k);
// tcgen05 pooled GEMM (stage A): approximate pooled scores -> per-label candidate rows + bound
int sdk_poolgemm_supported(int32_t Dp);
int sdk_launch_poolgemm_candidates(sdk_ctx* c, const __nv_bfloat16* d_bank, int64_t P,
                                   const __nv_bfloat16* d_seg, int64_t N, int32_t Dp,
                                   const int64_t* d_goff, int32_t G, int32_t pool, float tau,
                                   int32_t ncand, int32_t* d_cand_row /*[G,ncand]*/,
                                   float* d_gbound /*[G]*/);
int sdk_launch_probe_read(sdk_ctx* c);     // one plain read pass over the bf16 bank operands (bench.py's practical HBM ceiling)
// <= 8 query segments: the small-query latency path (gemv.cu) -- label offsets, normalise, HBM-bound bank stream and
// per-CTA top lists in one kernel; merge, canonical re-score, select and certificate in a second; nothing synchronised.
// Failed certificates are left on c->fb_list / flags[SDK_FLAG_FB] for sdk_settle.
int sdk_gemv_applicable(int64_t N, int32_t Dp);
int sdk_launch_gemv_identify(sdk_ctx* c, const void* d_seg_raw, int32_t in_dtype, const int32_t* d_seg_label, int32_t label_base, int64_t N,
                             int32_t L, int32_t pool, double threshold, int32_t k, float tau, float eps, int32_t ncand, int32_t* d_flags,
                             int64_t* o_row, float* o_score, int32_t* o_count, uint8_t* o_trust, int32_t* o_spk);
int sdk_launch_poolgemm_remerge(sdk_ctx* c, const int64_t* d_goff, const int32_t* d_glist, int32_t ngroups, float tau,
                                int32_t ncand, int32_t* d_cand_row, float* d_gbound);
// tcgen05 accumulate-pooling GEMM (mean pooling, >= 128 label groups): normalises the RAW segments into the group-
// interleaved bf16 layout (c->seg_bf16), pools inside the MMA accumulation; returns the row addressing of the layout
int sdk_poolacc_applicable(int32_t Dp, int32_t G, int32_t pool);
// plan of the layout for G label groups / N segments scored against P rows; *steps_out = 256-row steps (one stream sync)
int sdk_poolacc_plan(sdk_ctx* c, const int64_t* d_goff, int32_t G, int64_t N, int64_t P, int32_t Dp, int64_t* steps_out);
// mode 0: candidates (+ merge) into d_cand_row / d_gbound; mode 1: dense out[row, g] (config 5).  The interleaved
// operands are written to `il` (grown as needed); *d_grp_out = per-group addressing for the canonical re-score.
int sdk_launch_poolacc(sdk_ctx* c, const void* d_seg_raw, int32_t in_dtype, const int32_t* d_seg_label, int32_t label_base, int64_t N, int32_t D,
                       int32_t Dp, const __nv_bfloat16* d_rows, int64_t P, const int64_t* d_goff, int32_t G, int64_t S,
                       int32_t mode, float tau, int32_t ncand, int32_t* d_cand_row, float* d_gbound, float* d_dense,
                       sdk_buf& il, const PaGroup** d_grp_out);
// stage A over an already interleaved matrix whose plan is in c->pa_* (pool-first centroids)
int sdk_launch_poolacc_prepared(sdk_ctx* c, const void* d_il, int64_t n_rows, int32_t Dp, const __nv_bfloat16* d_rows, int64_t P,
                                const int64_t* d_goff, int32_t G, float tau, int32_t ncand, int32_t* d_cand_row, float* d_gbound);
// tcgen05 pooled GEMM, dense output out[row, g] (config 5)
int sdk_launch_poolgemm_dense(sdk_ctx* c, const __nv_bfloat16* d_rows, int64_t P,
                              const __nv_bfloat16* d_cols, int64_t N, int32_t Dp,
                              const int64_t* d_goff, int32_t G, int32_t pool, float* d_out /*[P,G]*/);

// ---- device helpers ---------------------------------------------------------------------------
// Raw embedding rows as the caller stores them: fp32, or IEEE fp16 (half the PCIe / HBM bytes; widening to fp32 is
// exact, so the canonical arithmetic downstream is unchanged).  ld4 = elements 4q .. 4q+3 of a row, streaming load.
#define SDK_IN_F32 0
#define SDK_IN_F16 1
template <typename T> struct sdk_in;
template <> struct sdk_in<float> {
    static constexpr int align = 16;
    static __device__ __forceinline__ float4 ld4(const float* row, int q) { return __ldcs(reinterpret_cast<const float4*>(row) + q); }
    static __device__ __forceinline__ float ld1(const float* row, int e) { return row[e]; }
};
template <> struct sdk_in<__half> {
    static constexpr int align = 8;
    static __device__ __forceinline__ float4 ld4(const __half* row, int q) {
        const uint2 u = __ldcs(reinterpret_cast<const uint2*>(row) + q);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float ld1(const __half* row, int e) { return __half2float(row[e]); }
};
static inline size_t sdk_in_size(int32_t in_dtype) { return in_dtype == SDK_IN_F16 ? 2 : 4; }
__device__ __forceinline__ uint32_t sdk_fkey(float f) {   // order-preserving float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sdk_funkey(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ float sdk_pool_finish(long long q, long long n, int pool) {
    if (n <= 0) return 0.0f;
    if (pool == 0) return (float)((double)q / ((double)n * SDK_Q30));
    return (float)((double)q / SDK_Q30);
}
