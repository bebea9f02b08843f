// api.cu -- the extern "C" boundary declared in include/sdk_b200.h.
#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <mutex>

#include "common.cuh"

static std::string g_last_error;
static std::mutex g_err_mu;

int sdk_fail(sdk_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_error = msg;
    return code;
}

int sdk_reserve(sdk_ctx* c, sdk_buf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return SDK_OK;
    if (b.p) { cudaStreamSynchronize(c->stream); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return sdk_fail(c, SDK_ENOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
    }
    b.cap = want;
    return SDK_OK;
}

static void sdk_release(sdk_buf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// ---- NCCL, loaded lazily so that single-GPU use has no link-time dependency ---------------------
typedef struct { char internal[128]; } sdk_nccl_uid_t;
struct sdk_nccl_api {
    void* h = nullptr;
    int (*GetUniqueId)(sdk_nccl_uid_t*) = nullptr;
    int (*CommInitRank)(void**, int, sdk_nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*CommAbort)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static sdk_nccl_api g_nccl;
static int sdk_nccl_load() {
    if (g_nccl.h) return SDK_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return sdk_fail(nullptr, SDK_ENCCL, std::string("dlopen libnccl failed: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(sdk_nccl_uid_t*))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, sdk_nccl_uid_t, int))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.CommAbort = (int (*)(void*))dlsym(g_nccl.h, "ncclCommAbort");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.h, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather)
        return sdk_fail(nullptr, SDK_ENCCL, "libnccl is missing required symbols");
    return SDK_OK;
}
#define SDK_NCCL_CHAR 0  /* ncclChar / ncclInt8 */

extern "C" {

int sdk_abi_version(void) { return SDK_ABI_VERSION; }

int sdk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int sdk_nccl_unique_id(void* out128) {
    if (!out128) return sdk_fail(nullptr, SDK_EINVAL, "out128 is NULL");
    SDK_TRY(sdk_nccl_load());
    sdk_nccl_uid_t id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return sdk_fail(nullptr, SDK_ENCCL, "ncclGetUniqueId failed");
    memcpy(out128, &id, 128);
    return SDK_OK;
}

int sdk_create(sdk_ctx** out, int device, int world, int rank, const void* nccl_uid) {
    if (!out) return sdk_fail(nullptr, SDK_EINVAL, "out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return sdk_fail(nullptr, SDK_EINVAL, "bad world/rank");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return sdk_fail(nullptr, SDK_ENODEV, "no CUDA device visible (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n) return sdk_fail(nullptr, SDK_ENODEV, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return sdk_fail(nullptr, SDK_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return sdk_fail(nullptr, SDK_ENODEV, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                                 ", this library is built for sm_100a only");
    sdk_ctx* c = new sdk_ctx();
    c->device = device;
    c->world = world;
    c->rank = rank;
    c->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return sdk_fail(nullptr, SDK_ECUDA, "cudaSetDevice/cudaStreamCreate failed");
    }
    cudaEventCreate(&c->ev_t0);
    cudaEventCreate(&c->ev_t1);
    // cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency)
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        c->tmap_encode = fn;
    else
        cudaGetLastError();
    if (world > 1) {
        if (!nccl_uid) { sdk_destroy(c); return sdk_fail(nullptr, SDK_EINVAL, "world > 1 needs an NCCL unique id"); }
        int r = sdk_nccl_load();
        if (r != SDK_OK) { sdk_destroy(c); return r; }
        sdk_nccl_uid_t id;
        memcpy(&id, nccl_uid, 128);
        int e = g_nccl.CommInitRank(&c->nccl_comm, world, id, rank);
        if (e != 0) {
            std::string m = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
            sdk_destroy(c);
            return sdk_fail(nullptr, SDK_ENCCL, m);
        }
    }
    *out = c;
    return SDK_OK;
}

void sdk_destroy(sdk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    sdk_buf* bufs[] = {&c->bank_f32, &c->bank_bf16, &c->row_speaker, &c->row_trust, &c->seg_raw, &c->seg_lab,
                       &c->seg_f32, &c->seg_bf16, &c->goff, &c->qpool, &c->dense, &c->flags, &c->cand_row,
                       &c->cand_val, &c->cand_cnt, &c->gbound, &c->slot_cnt, &c->slot_row, &c->slot_val,
                       &c->slot_bound, &c->range_g, &c->kth, &c->fb_list, &c->fb_rows, &c->out_pack, &c->gather, &c->stage_seg[0], &c->stage_seg[1], &c->stage_seg[2],
                       &c->stage_lab[0], &c->stage_lab[1], &c->stage_lab[2], &c->pa_hist, &c->pa_sorted, &c->pa_pos, &c->pa_col_group,
                       &c->pa_col_meta, &c->pa_blockT, &c->pa_step0, &c->pa_grp, &c->pa_col_last, &c->seg_il, &c->fb_list2, &c->cand_row2, &c->qpool2,
                       &c->cent_sum, &c->cent_seg, &c->goff2};
    for (sdk_buf* b : bufs) sdk_release(*b);
    if (c->h_pack) cudaFreeHost(c->h_pack);
    for (auto& p : c->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (int b = 0; b < SDK_NSTAGE; ++b) {
        if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]);
        if (c->ev_consumed[b]) cudaEventDestroy(c->ev_consumed[b]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* sdk_last_error(sdk_ctx* c) {
    if (c) return c->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mu);
    return g_last_error.c_str();
}

int sdk_set_option(sdk_ctx* c, const char* key, double value) {
    if (!c || !key) return sdk_fail(c, SDK_EINVAL, "NULL ctx/key");
    std::string k(key);
    if (k == "path") {
        if (value < 0 || value > 2) return sdk_fail(c, SDK_EINVAL, "path must be 0, 1 or 2");
        c->opt_path = (int)value;
    } else if (k == "eps") c->opt_eps = value;
    else if (k == "profile") c->opt_profile = value != 0;
    else if (k == "cand") {
        if (value < 1 || value > 64) return sdk_fail(c, SDK_EINVAL, "cand must be in 1..64");
        c->opt_cand = (int)value;
    } else if (k == "cta_group") {
        if (value != 0 && value != 1 && value != 2) return sdk_fail(c, SDK_EINVAL, "cta_group must be 0 (auto), 1 or 2");
        c->opt_cta_group = (int)value;
    } else if (k == "acc") {
        if (value != 0 && value != 1 && value != 2) return sdk_fail(c, SDK_EINVAL, "acc must be 0 (off), 1 (auto) or 2 (force)");
        c->opt_acc = (int)value;
    } else if (k == "gemv") {
        if (value != 0 && value != 1) return sdk_fail(c, SDK_EINVAL, "gemv must be 0 or 1");
        c->opt_gemv = (int)value;
    } else if (k == "chunk_mb") {
        if (value < 1 || value > 65536) return sdk_fail(c, SDK_EINVAL, "chunk_mb must be in 1..65536");
        c->opt_chunk_mb = (int)value;
    } else if (k == "kth") {
        if (value != 0 && value != 1 && value != 2) return sdk_fail(c, SDK_EINVAL, "kth must be 0 (off), 1 (auto) or 2 (on)");
        c->opt_kth = (int)value;
    } else if (k == "poolfirst") {
        if (value != 0 && value != 1) return sdk_fail(c, SDK_EINVAL, "poolfirst must be 0 or 1");
        c->opt_poolfirst = (int)value;
    } else if (k == "inject_fail") {
        c->opt_inject_fail = value != 0;          // test knob: the next local identify pass fails after its first kernels
    } else return sdk_fail(c, SDK_EINVAL, "unknown option " + k);
    return SDK_OK;
}

// ---- bank ---------------------------------------------------------------------------------------
static int sdk_check_bank_args(sdk_ctx* c, const void* rows, const void* spk, int64_t P, int32_t D, int32_t dtype) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    // an EMPTY shard (P == 0) is legal in the row-sharded mode: the rank still joins every collective with empty lists
    if (P < (c->world > 1 ? 0 : 1) || D < 1 || D > 8192) return sdk_fail(c, SDK_EINVAL, "bank needs P >= 1 and 1 <= D <= 8192");
    if (P > 0x7fffffffLL) return sdk_fail(c, SDK_EINVAL, "at most 2^31-1 rows per shard");
    if (P > 0 && (!rows || !spk)) return sdk_fail(c, SDK_EINVAL, "rows/row_speaker is NULL");
    if (dtype != SDK_DTYPE_F32 && dtype != SDK_DTYPE_BF16) return sdk_fail(c, SDK_EINVAL, "dtype must be 0 (fp32) or 1 (bf16)");
    return SDK_OK;
}

int sdk_bank_load_dev(sdk_ctx* c, const float* d_rows, const int32_t* d_row_speaker, const uint8_t* d_row_trust,
                      int64_t P, int32_t D, int32_t dtype, int64_t global_row_offset) {
    SDK_TRY(sdk_check_bank_args(c, d_rows, d_row_speaker, P, D, dtype));
    cudaSetDevice(c->device);
    c->P = 0;
    c->bank_ok = false;
    int32_t Dp = (D + 63) / 64 * 64;
    if (P == 0) {
        c->D = D; c->Dp = Dp; c->dtype = dtype; c->row_offset = global_row_offset;
        c->bank_ok = true;
        c->have_results = false;
        return SDK_OK;
    }
    SDK_TRY(sdk_reserve(c, c->bank_bf16, (size_t)P * Dp * 2));
    if (dtype == SDK_DTYPE_F32) SDK_TRY(sdk_reserve(c, c->bank_f32, (size_t)P * D * 4));
    SDK_TRY(sdk_reserve(c, c->row_speaker, (size_t)P * 4));
    SDK_TRY(sdk_reserve(c, c->row_trust, (size_t)P));
    SDK_CUDA(c, cudaMemcpyAsync(c->row_speaker.p, d_row_speaker, (size_t)P * 4, cudaMemcpyDeviceToDevice, c->stream));
    if (d_row_trust) SDK_CUDA(c, cudaMemcpyAsync(c->row_trust.p, d_row_trust, (size_t)P, cudaMemcpyDeviceToDevice, c->stream));
    else SDK_CUDA(c, cudaMemsetAsync(c->row_trust.p, SDK_TRUST_UNKNOWN, (size_t)P, c->stream));
    SDK_TRY(sdk_launch_normalize(c, d_rows, P, D, Dp, dtype == SDK_DTYPE_F32 ? (float*)c->bank_f32.p : nullptr,
                                 (__nv_bfloat16*)c->bank_bf16.p));
    c->P = P; c->D = D; c->Dp = Dp; c->dtype = dtype; c->row_offset = global_row_offset;
    c->bank_ok = true;
    c->have_results = false;
    return SDK_OK;
}

int sdk_bank_load(sdk_ctx* c, const float* rows, const int32_t* row_speaker, const uint8_t* row_trust, int64_t P,
                  int32_t D, int32_t dtype, int64_t global_row_offset) {
    SDK_TRY(sdk_check_bank_args(c, rows, row_speaker, P, D, dtype));
    if (P == 0) return sdk_bank_load_dev(c, nullptr, nullptr, nullptr, 0, D, dtype, global_row_offset);
    // rows of one speaker must be contiguous (the select kernel and the shard cut rely on it)
    {
        std::vector<int32_t> seen;
        int32_t prev = row_speaker[0];
        if (prev < 0) return sdk_fail(c, SDK_EINVAL, "row_speaker must be >= 0");
        seen.push_back(prev);
        for (int64_t i = 1; i < P; ++i) {
            int32_t s = row_speaker[i];
            if (s < 0) return sdk_fail(c, SDK_EINVAL, "row_speaker must be >= 0");
            if (s != prev) { seen.push_back(s); prev = s; }
        }
        std::sort(seen.begin(), seen.end());
        if (std::adjacent_find(seen.begin(), seen.end()) != seen.end())
            return sdk_fail(c, SDK_EINVAL, "rows of one speaker must be contiguous in the bank");
    }
    cudaSetDevice(c->device);
    SDK_TRY(sdk_reserve(c, c->seg_raw, (size_t)P * D * 4));
    SDK_TRY(sdk_reserve(c, c->seg_lab, (size_t)P * 4 + (size_t)P));
    SDK_CUDA(c, cudaMemcpyAsync(c->seg_raw.p, rows, (size_t)P * D * 4, cudaMemcpyHostToDevice, c->stream));
    SDK_CUDA(c, cudaMemcpyAsync(c->seg_lab.p, row_speaker, (size_t)P * 4, cudaMemcpyHostToDevice, c->stream));
    uint8_t* d_tr = nullptr;
    if (row_trust) {
        for (int64_t i = 0; i < P; ++i)
            if (row_trust[i] > SDK_TRUST_UNKNOWN) return sdk_fail(c, SDK_EINVAL, "row_trust code out of range");
        d_tr = (uint8_t*)c->seg_lab.p + (size_t)P * 4;
        SDK_CUDA(c, cudaMemcpyAsync(d_tr, row_trust, (size_t)P, cudaMemcpyHostToDevice, c->stream));
    }
    int r = sdk_bank_load_dev(c, (const float*)c->seg_raw.p, (const int32_t*)c->seg_lab.p, d_tr, P, D, dtype, global_row_offset);
    if (r != SDK_OK) return r;
    SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    return SDK_OK;
}

// ---- identify -----------------------------------------------------------------------------------
static int sdk_flag_error(sdk_ctx* c, const int32_t* f) {
    if (f[SDK_FLAG_LABEL] & 1) return sdk_fail(c, SDK_EINVAL, "seg_label out of range [0,L)");
    if (f[SDK_FLAG_LABEL] & 2) return sdk_fail(c, SDK_EINVAL, "seg_label must be non-decreasing (segments sorted by label group)");
    if (f[SDK_FLAG_PEER] != 0)
        return sdk_fail(c, SDK_EPEER, "rank " + std::to_string(f[SDK_FLAG_PEER] - 1) + " of the row-sharded bank failed in this identify call "
                                      "(bad labels or a local error); the merged result is not valid");
    return SDK_OK;
}
static int sdk_check_flags(sdk_ctx* c, const int32_t* d_flags) {   // after a stream sync
    int32_t f[SDK_NFLAGS] = {0};
    SDK_CUDA(c, cudaMemcpy(f, d_flags, sizeof(f), cudaMemcpyDeviceToHost));
    return sdk_flag_error(c, f);
}

static int sdk_allgather_merge(sdk_ctx* c, int32_t L, int32_t k);
static int sdk_identify_core(sdk_ctx* c, const void* d_seg, const int32_t* d_seg_label, int64_t N, int32_t L,
                             int32_t label_base, int32_t pool, double threshold, int32_t k, bool allow_lazy);
static int sdk_settle(sdk_ctx* c, const int32_t* known_flags = nullptr);

static int sdk_check_identify_args(sdk_ctx* c, const void* seg, const void* lab, int64_t N, int32_t L, int32_t pool,
                                   double threshold, int32_t k) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    if (!c->bank_ok) return sdk_fail(c, SDK_ESTATE, "sdk_identify before sdk_bank_load");
    if (k < 1 || k > SDK_MAX_K) return sdk_fail(c, SDK_EINVAL, "k must be in 1..32");
    if (L < 1 || N < 0) return sdk_fail(c, SDK_EINVAL, "need L >= 1 and N >= 0");
    if (pool != SDK_POOL_MEAN && pool != SDK_POOL_MAX) return sdk_fail(c, SDK_EINVAL, "pool must be 0 (mean) or 1 (max)");
    if (N > 0 && (!seg || !lab)) return sdk_fail(c, SDK_EINVAL, "seg/seg_label is NULL");
    if (!(threshold == threshold)) return sdk_fail(c, SDK_EINVAL, "threshold is NaN");
    return SDK_OK;
}

static size_t sdk_al16(size_t x) { return (x + 15) / 16 * 16; }
// Lays the result record of L label groups x k matches out in c->out_pack (see sdk_out_view) and clears its flags.
static int sdk_reserve_results(sdk_ctx* c, int32_t L, int32_t k) {
    sdk_out_view v;
    const size_t n = (size_t)L * k;
    size_t o = SDK_NFLAGS * 4;
    v.off_row = o;      o = sdk_al16(o + n * 8);
    v.off_score = o;    o = sdk_al16(o + n * 4);
    v.off_spk = o;      o = sdk_al16(o + n * 4);
    v.off_count = o;    o = sdk_al16(o + (size_t)L * 4);
    v.off_trust = o;    o = sdk_al16(o + n);
    v.gather_bytes = o;
    v.off_as_score = o; o = sdk_al16(o + (size_t)L * 8);
    v.off_as_cscore = o; o = sdk_al16(o + (size_t)L * 24);
    v.off_as_idx = o;   o = sdk_al16(o + (size_t)L * 4);
    v.off_as_conf = o;  o = sdk_al16(o + (size_t)L * 4);
    v.off_as_cidx = o;  o = sdk_al16(o + (size_t)L * 12);
    v.bytes = o;
    SDK_TRY(sdk_reserve(c, c->out_pack, v.bytes));
    char* b = (char*)c->out_pack.p;
    v.flags = (int32_t*)b;
    v.row = (int64_t*)(b + v.off_row);
    v.score = (float*)(b + v.off_score);
    v.spk = (int32_t*)(b + v.off_spk);
    v.count = (int32_t*)(b + v.off_count);
    v.trust = (uint8_t*)(b + v.off_trust);
    v.as_score = (double*)(b + v.off_as_score);
    v.as_cscore = (double*)(b + v.off_as_cscore);
    v.as_idx = (int32_t*)(b + v.off_as_idx);
    v.as_conf = (int32_t*)(b + v.off_as_conf);
    v.as_cidx = (int32_t*)(b + v.off_as_cidx);
    c->out = v;
    SDK_CUDA(c, cudaMemsetAsync(v.flags, 0, SDK_NFLAGS * 4, c->stream));
    return SDK_OK;
}

// One pass of the hot path over the label groups [label_base, label_base + L): segments d_seg (N rows) carry
// GLOBAL label ids; results land at group offset label_base of the context's result arrays.  The raw rows are fp32 or
// fp16 (c->in_dtype, set by the entry point).
// Label groups whose top-k certificate failed are re-done exhaustively in the canonical arithmetic (rows of `fb`).
static int sdk_exhaustive_fallback(sdk_ctx* c, const int32_t* fb, int32_t nfb, const void* seg_ops, const PaGroup* seg_grp, int32_t pool,
                                   double threshold, int32_t k, int64_t* o_row, float* o_score, int32_t* o_count, uint8_t* o_trust, int32_t* o_spk) {
    const bool bf16 = c->dtype == SDK_DTYPE_BF16;
    const int64_t P = c->P;
    const void* bank_ops = bf16 ? c->bank_bf16.p : c->bank_f32.p;
    const int32_t pitch = bf16 ? c->Dp : c->D;
    const int32_t chunk = (int32_t)std::max<int64_t>(1, std::min<int64_t>(nfb, (int64_t)(1u << 28) / std::max<int64_t>(P, 1)));
    for (int32_t done = 0; done < nfb; done += chunk) {
        int32_t m = std::min(chunk, nfb - done);
        SDK_TRY(sdk_reserve(c, c->dense, (size_t)m * P * 8));
        const int32_t* gl = fb + done;
        SDK_TRY(sdk_launch_exact(c, seg_ops, bank_ops, bf16, c->D, pitch, (const int64_t*)c->goff.p, gl, m, nullptr, P,
                                 pool, (long long*)c->dense.p, seg_grp));
        SDK_TRY(sdk_launch_select(c, (const long long*)c->dense.p, (const int64_t*)c->goff.p, gl, m, nullptr, P, pool,
                                  (const int32_t*)c->row_speaker.p, (const uint8_t*)c->row_trust.p, threshold, k,
                                  c->row_offset, nullptr, 0.f, 0.f, nullptr, 0, nullptr, nullptr, o_row, o_score, o_count, o_trust, o_spk));
    }
    return SDK_OK;
}

// The small-query path (gemv.cu) does not stop for its certificate: the identify call returns with the kernels queued,
// and whoever looks at the results next (fetch, last_path, stage_a, the all-gather) settles the open questions first --
// the label flags, and the exhaustive pass for the (rare) labels on the fall-back list.  `known_flags`: the flag words
// if the caller has just read them (the small-record fetch reads the whole record in one copy).
static int sdk_settle(sdk_ctx* c, const int32_t* known_flags) {
    if (!c->lazy.active) return SDK_OK;
    c->lazy.active = false;
    int32_t hf[2] = {0, 0};
    if (known_flags) { hf[0] = known_flags[SDK_FLAG_LABEL]; hf[1] = known_flags[SDK_FLAG_FB]; }
    else {
        SDK_CUDA(c, cudaMemcpyAsync(hf, c->out.flags, 8, cudaMemcpyDeviceToHost, c->stream));
        SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (hf[0] & 1) return sdk_fail(c, SDK_EINVAL, "seg_label out of range [0,L)");
    if (hf[0] & 2) return sdk_fail(c, SDK_EINVAL, "seg_label must be non-decreasing (segments sorted by label group)");
    const int32_t nfb = hf[1];
    if (nfb <= 0) return SDK_OK;
    const sdk_lazy& z = c->lazy;
    c->last_fallback += nfb;
    SDK_TRY(sdk_exhaustive_fallback(c, (const int32_t*)c->fb_list.p, nfb, z.seg_ops, nullptr, z.pool, z.threshold, z.k, z.o_row, z.o_score,
                                    z.o_count, z.o_trust, z.o_spk));
    SDK_CUDA(c, cudaMemsetAsync(c->out.flags + SDK_FLAG_FB, 0, 4, c->stream));
    if (c->have_assign) {                                  // the assignment was computed from the unsettled lists: redo it
        const sdk_out_view& v = c->out;
        SDK_TRY(sdk_launch_assign(c, v.row, v.score, v.trust, v.count, c->L, c->k, c->assign_thr, c->assign_min_trust, v.as_idx,
                                  v.as_score, v.as_conf, v.as_cidx, v.as_cscore));
    }
    return SDK_OK;
}

static int sdk_identify_core(sdk_ctx* c, const void* d_seg, const int32_t* d_seg_label, int64_t N, int32_t L,
                             int32_t label_base, int32_t pool, double threshold, int32_t k, bool allow_lazy) {
    const int32_t D = c->D, Dp = c->Dp;
    const int64_t P = c->P;
    const bool bf16 = c->dtype == SDK_DTYPE_BF16;
    int64_t* o_row = c->out.row + (size_t)label_base * k;
    float* o_score = c->out.score + (size_t)label_base * k;
    int32_t* o_count = c->out.count + label_base;
    uint8_t* o_trust = c->out.trust + (size_t)label_base * k;
    int32_t* o_spk = c->out.spk + (size_t)label_base * k;

    int32_t* d_flags = c->out.flags;
    // path choice: tcgen05 only where the contraction is big enough to be a real dense GEMM
    const double macs = (double)N * (double)P * (double)Dp;
    int path = c->opt_path;
    // (the exact kernel is thread-per-segment: label groups of a few segments against a big bank leave most of a CTA idle,
    //  so those shapes -- e.g. 8 pooled centroids against a 125k-row shard -- go to the tensor path much earlier)
    if (path == 0) {
        const bool big = macs > 536870912.0 || (macs > 67108864.0 && (double)N < 64.0 * (double)L);
        path = (big && sdk_poolgemm_supported(Dp) && c->tmap_encode) ? 2 : 1;
    }
    if (path == 2 && !(sdk_poolgemm_supported(Dp) && c->tmap_encode))
        return sdk_fail(c, SDK_EINVAL, "tcgen05 path not available for this D / driver");
    c->last_path = path;

    // accumulate-pooling (poolacc.cu): mean pooling is done inside the MMA accumulation; it normalises the raw segments
    // straight into its group-interleaved bf16 layout.  Auto mode takes it where the generic kernel is bound by its
    // epilogue (D <= 256) or where there are enough groups for tight size classes, and only if the layout's zero
    // padding stays under 12 %.
    // a handful of query segments: one HBM-bound pass over the bank on the CUDA cores (gemv.cu)
    // (its candidate slots are indexed by label id: a handful of segments scattered over very many labels stays generic)
    const bool use_gemv = path == 2 && c->opt_gemv && sdk_gemv_applicable(N, Dp) && L <= 64;
    if (use_gemv && !c->opt_inject_fail) {
        // the small-query latency path: two kernels (label offsets, normalise, bank stream, per-CTA top lists | merge,
        // canonical re-score, select, certificate), no host round trip inside the call
        c->last_path = 4;
        const float kc = (float)(Dp / 16);
        float eps = (bf16 ? 0.f : 1.2e-2f) + 1.05f * 7.15255737e-07f * kc + 4.8e-7f;      // margin model: see the tensor path below
        if (c->opt_eps >= 0) eps = (float)c->opt_eps;
        int ncand = std::min(64, std::max(c->opt_cand, k + 6));
        float tau = (float)(threshold - 2.0 * (double)eps);
        if (!(tau > -3.0e38f)) tau = -3.0e38f;
        c->last_eps_base = eps;
        c->last_eps_chain = 0.f;
        c->last_ncand = ncand;
        c->last_cand_groups = L;
        c->kth_on = false;
        SDK_TRY(sdk_launch_gemv_identify(c, d_seg, c->in_dtype, d_seg_label, label_base, N, L, pool, threshold, k, tau, eps, ncand, d_flags,
                                         o_row, o_score, o_count, o_trust, o_spk));
        sdk_lazy& z = c->lazy;
        z.active = true;
        z.seg_ops = bf16 ? c->seg_bf16.p : c->seg_f32.p;
        z.pool = pool; z.threshold = threshold; z.k = k;
        z.o_row = o_row; z.o_score = o_score; z.o_count = o_count; z.o_trust = o_trust; z.o_spk = o_spk;
        if (!allow_lazy) SDK_TRY(sdk_settle(c));
        return SDK_OK;
    }
    SDK_TRY(sdk_reserve(c, c->goff, (size_t)(L + 1) * 8));
    SDK_TRY(sdk_launch_group_offsets(c, d_seg_label, N, L, label_base, (int64_t*)c->goff.p, d_flags));
    if (c->opt_inject_fail) {
        c->opt_inject_fail = 0;
        return sdk_fail(c, SDK_EINVAL, "injected failure (option inject_fail)");
    }
    bool use_acc = path == 2 && c->opt_acc && sdk_poolacc_applicable(Dp, L, pool) && D % 4 == 0 && D <= 2048 &&
                   (uintptr_t)d_seg % (c->in_dtype == SDK_IN_F16 ? 8 : 16) == 0;     // its scatter-normalise reads the raw rows with 128-bit (fp16: 64-bit) loads
    if (use_acc && c->opt_acc != 2 && !(Dp <= 256 || L > 2048)) use_acc = false;
    int64_t acc_steps = 0;
    if (path == 2) {   // the plan of either tcgen05 kernel trusts goff: reject bad labels before going on
        int32_t lf = 0;
        SDK_CUDA(c, cudaMemcpyAsync(&lf, d_flags, 4, cudaMemcpyDeviceToHost, c->stream));
        SDK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (lf & 1) return sdk_fail(c, SDK_EINVAL, "seg_label out of range [0,L)");
        if (lf & 2) return sdk_fail(c, SDK_EINVAL, "seg_label must be non-decreasing (segments sorted by label group)");
    }
    // pool-first (option "poolfirst", mean pooling): stage A contracts the label centroids instead of the segments -- a
    // different algorithm (N/L times fewer flops, HBM-bound on K1); stage B and the certificate are unchanged
    const bool use_pf = path == 2 && c->opt_poolfirst && pool == SDK_POOL_MEAN && N > 0 && sdk_poolfirst_applicable(d_seg, c->in_dtype, D, Dp);
    if (use_pf) use_acc = false;
    if (use_acc) {
        SDK_TRY(sdk_poolacc_plan(c, (const int64_t*)c->goff.p, L, N, P, Dp, &acc_steps));
        if (c->opt_acc != 2 && (double)acc_steps * 256.0 > 1.12 * (double)N + 1024.0) use_acc = false;   // acc == 2 forces it (tests)
    }
    if (use_acc) c->last_path = 3;
    if (use_pf) c->last_path = 5;
    const bool need_bf16 = (bf16 || path == 2) && !use_acc;
    if (!bf16) SDK_TRY(sdk_reserve(c, c->seg_f32, (size_t)N * D * 4));
    if (need_bf16) SDK_TRY(sdk_reserve(c, c->seg_bf16, (size_t)N * Dp * 2));
    int64_t pf_rows = 0;
    if (use_pf) {
        SDK_TRY(sdk_launch_normalize_centroid(c, d_seg, c->in_dtype, d_seg_label, label_base, N, D, Dp, bf16 ? nullptr : (float*)c->seg_f32.p,
                                              (__nv_bfloat16*)c->seg_bf16.p, (const int64_t*)c->goff.p, L, bf16 ? 1 : 0, &pf_rows));
    } else if (!bf16 || need_bf16)
        SDK_TRY(sdk_launch_normalize_in(c, d_seg, c->in_dtype, N, D, Dp, bf16 ? nullptr : (float*)c->seg_f32.p,
                                        need_bf16 ? (__nv_bfloat16*)c->seg_bf16.p : nullptr));
    const void* bank_ops = bf16 ? c->bank_bf16.p : c->bank_f32.p;
    const int32_t pitch = bf16 ? Dp : D;
    const PaGroup* seg_grp = nullptr;      // row addressing of the segment operands (identity unless interleaved)

    if (path == 1) {
        c->last_ncand = 0;
        const void* seg_ops = bf16 ? c->seg_bf16.p : c->seg_f32.p;
        SDK_TRY(sdk_reserve(c, c->qpool, (size_t)L * P * 8));
        SDK_TRY(sdk_launch_exact(c, seg_ops, bank_ops, bf16, D, pitch, (const int64_t*)c->goff.p, nullptr, L, nullptr, P,
                                 pool, (long long*)c->qpool.p, nullptr, N));
        SDK_TRY(sdk_launch_select(c, (const long long*)c->qpool.p, (const int64_t*)c->goff.p, nullptr, L, nullptr, P, pool,
                                  (const int32_t*)c->row_speaker.p, (const uint8_t*)c->row_trust.p, threshold, k,
                                  c->row_offset, nullptr, 0.f, 0.f, nullptr, 0, nullptr, nullptr, o_row, o_score, o_count, o_trust, o_spk));
    } else {
        // stage A: tcgen05 pooled GEMM -> per-label candidate rows + bound on everything else.
        // Certificate margin: eps_g = eps_base + eps_chain * chain_g bounds |stage-A pooled score - canonical| of group g
        // (select.cu; derivation in DESIGN.md section 2).  u = 6 * 2^-23 per tensor-core accumulator update (one K = 16 MMA),
        // relative to the running magnitude: the accumulation TRUNCATES -- measured on all-positive data (every partial
        // sum grows, tests/test_gpu_certificate.py) the loss is a systematic ~2.2 ulp per update, i.e. about five
        // truncated addends of up to one ulp each; 6 ulp is the worst case of that model.  Normalised operands, so
        // |dot| <= ~1.01 and a partial pooled sum of t segments is <= ~1.01 t.
        //   one segment's dot product: Dp/16 updates of magnitude <= 1             -> eps_dot = u * Dp/16
        //   accumulate-pooling (chain = T segments per column): sum_t u*(Dp/16)*t / n  -> eps_chain = u/2 * Dp/16 per segment
        //   epilogue mean pooling: fp32 adds of 32-column block sums, RN            -> 2^-24 per block (chain = n/32 + 70)
        //   fp32 bank: stage A rounds both operands to bf16 (2 * 2^-9 relative)     -> + 2^-8 * 1.02, covered by 1.2e-2
        const float u_mma = 7.15255737e-07f;                     // 6 * 2^-23
        const float kc = (float)(Dp / 16);
        float eps_base = (bf16 ? 0.f : 1.2e-2f) + 1.05f * u_mma * kc + 4.8e-7f, eps_chain = 0.f;
        int32_t chain_div = 0;
        double chain_max = 0.0;
        if (use_acc) { eps_chain = 0.5f * 1.05f * u_mma * kc; chain_max = (double)c->pa_chain_max; }
        else if (use_pf) {
            // pool-first: |<hi + lo, b> computed - canonical pooled score| <=  two dot products at operand magnitude 2 (the
            // doubled halves) accumulated in one tile (+ slack 4.3e-6), the hi/lo split residual (2^-18), the fp32
            // centroid (64 register adds: 64 * 2^-24; n/64 atomic adds: 2^-24/64 per segment -> eps_chain), Q30 rounding
            eps_base = (bf16 ? 0.f : 1.2e-2f) + 2.1f * u_mma * kc + 4.3e-6f + 3.9e-6f + 4.1e-6f + 4.8e-7f;
            eps_chain = 9.4e-10f;
            chain_div = 1;
            chain_max = (double)(N + 70);
        }
        else if (pool == SDK_POOL_MEAN) { eps_chain = 6.1e-8f; chain_div = 32; chain_max = (double)(N / 32 + 70); }
        if (c->opt_eps >= 0) { eps_base = (float)c->opt_eps; eps_chain = 0.f; }       // test knob: fixed margin
        const float eps = eps_base + eps_chain * (float)chain_max;                     // launch-wide (candidate threshold tau)
        c->last_eps_base = eps_base;
        c->last_eps_chain = eps_chain;
        int ncand = std::max(c->opt_cand, k + 6);
        if (ncand > 64) ncand = 64;
        float tau = (float)(threshold - 2.0 * (double)eps);
        if (!(tau > -3.0e38f)) tau = -3.0e38f;
        SDK_TRY(sdk_reserve(c, c->cand_val, (size_t)L * ncand * 4));
        // low / no threshold: most bank rows pass tau, so the flushes prune against a running per-label bound on the
        // 64th best score instead of writing half the bank into the candidate slots (option "kth": 0 off, 1 auto, 2 on)
        c->kth_on = c->opt_kth == 2 || (c->opt_kth == 1 && tau < 0.25f);
        c->last_ncand = ncand;
        c->last_cand_groups = L;
        SDK_TRY(sdk_reserve(c, c->cand_row, (size_t)L * ncand * 4));
        SDK_TRY(sdk_reserve(c, c->gbound, (size_t)L * 4));
        SDK_TRY(sdk_reserve(c, c->fb_list, (size_t)L * 4));
        const PaGroup* acc_grp = nullptr;       // accumulation chain per column (certificate margin of giant groups)
        c->slot_g0 = c->slot_g1 = 0;
        if (use_acc) {
            const PaGroup* ig = nullptr;
            SDK_TRY(sdk_launch_poolacc(c, d_seg, c->in_dtype, d_seg_label, label_base, N, D, Dp, (const __nv_bfloat16*)c->bank_bf16.p, P,
                                       (const int64_t*)c->goff.p, L, acc_steps, 0, tau, ncand, (int32_t*)c->cand_row.p,
                                       (float*)c->gbound.p, nullptr, c->seg_bf16, &ig));
            if (bf16) seg_grp = ig;                            // bf16 operands live in the interleaved matrix
            acc_grp = ig;
        } else if (use_pf) {
            // the centroid GEMM runs on the accumulate-pooling kernel: one column per label, two steps (hi, lo) per block
            SDK_TRY(sdk_launch_poolacc_prepared(c, c->cent_seg.p, pf_rows, Dp, (const __nv_bfloat16*)c->bank_bf16.p, P,
                                                (const int64_t*)c->goff2.p, L, tau, ncand, (int32_t*)c->cand_row.p, (float*)c->gbound.p));
        } else {
            SDK_TRY(sdk_launch_poolgemm_candidates(c, (const __nv_bfloat16*)c->bank_bf16.p, P, (const __nv_bfloat16*)c->seg_bf16.p,
                                                   N, Dp, (const int64_t*)c->goff.p, L, pool, tau, ncand,
                                                   (int32_t*)c->cand_row.p, (float*)c->gbound.p));
        }
        const void* seg_ops = bf16 ? c->seg_bf16.p : c->seg_f32.p;
        // stage B: canonical re-score of the candidates, ordered top-k, certificate
        SDK_TRY(sdk_reserve(c, c->qpool, (size_t)L * ncand * 8));
        SDK_TRY(sdk_launch_exact(c, seg_ops, bank_ops, bf16, D, pitch, (const int64_t*)c->goff.p, nullptr, L,
                                 (const int32_t*)c->cand_row.p, ncand, pool, (long long*)c->qpool.p, seg_grp));
        SDK_TRY(sdk_launch_select(c, (const long long*)c->qpool.p, (const int64_t*)c->goff.p, nullptr, L,
                                  (const int32_t*)c->cand_row.p, ncand, pool, (const int32_t*)c->row_speaker.p,
                                  (const uint8_t*)c->row_trust.p, threshold, k, c->row_offset, (const float*)c->gbound.p,
                                  eps_base, eps_chain, acc_grp, chain_div, d_flags + 1, (int32_t*)c->fb_list.p, o_row, o_score, o_count, o_trust, o_spk));
        // groups whose certificate failed are re-done exhaustively in the canonical arithmetic
        int32_t hf[2] = {0, 0};
        SDK_CUDA(c, cudaMemcpyAsync(hf, d_flags, 8, cudaMemcpyDeviceToHost, c->stream));
        SDK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (hf[0] & 1) return sdk_fail(c, SDK_EINVAL, "seg_label out of range [0,L)");
        if (hf[0] & 2) return sdk_fail(c, SDK_EINVAL, "seg_label must be non-decreasing (segments sorted by label group)");
        int32_t nfb = hf[1];
        const int32_t* fb = (const int32_t*)c->fb_list.p;
        // second chance (all groups' candidate slots still live): re-merge the failed groups with the
        // widest candidate list, re-score, certify again -- an exhaustive pass over a 125k-row shard costs ~2 ms per
        // group, this costs microseconds
        if (nfb > 0 && ncand < 64 && c->slot_g0 == 0 && c->slot_g1 == L) {
            const int32_t ncand2 = 64;
            SDK_TRY(sdk_reserve(c, c->cand_row2, (size_t)nfb * ncand2 * 4));
            SDK_TRY(sdk_reserve(c, c->qpool2, (size_t)nfb * ncand2 * 8));
            SDK_TRY(sdk_reserve(c, c->fb_list2, (size_t)L * 4));
            SDK_CUDA(c, cudaMemsetAsync(d_flags + 1, 0, 4, c->stream));
            SDK_TRY(sdk_launch_poolgemm_remerge(c, (const int64_t*)c->goff.p, fb, nfb, tau, ncand2, (int32_t*)c->cand_row2.p,
                                                (float*)c->gbound.p));
            SDK_TRY(sdk_launch_exact(c, seg_ops, bank_ops, bf16, D, pitch, (const int64_t*)c->goff.p, fb, nfb,
                                     (const int32_t*)c->cand_row2.p, ncand2, pool, (long long*)c->qpool2.p, seg_grp));
            SDK_TRY(sdk_launch_select(c, (const long long*)c->qpool2.p, (const int64_t*)c->goff.p, fb, nfb,
                                      (const int32_t*)c->cand_row2.p, ncand2, pool, (const int32_t*)c->row_speaker.p,
                                      (const uint8_t*)c->row_trust.p, threshold, k, c->row_offset, (const float*)c->gbound.p,
                                      eps_base, eps_chain, acc_grp, chain_div, d_flags + 1, (int32_t*)c->fb_list2.p, o_row, o_score, o_count, o_trust, o_spk));
            SDK_CUDA(c, cudaMemcpyAsync(hf, d_flags, 8, cudaMemcpyDeviceToHost, c->stream));
            SDK_CUDA(c, cudaStreamSynchronize(c->stream));
            c->last_retry += nfb;
            nfb = hf[1];
            fb = (const int32_t*)c->fb_list2.p;
        }
        c->last_fallback += nfb;
        SDK_TRY(sdk_exhaustive_fallback(c, fb, nfb, seg_ops, seg_grp, pool, threshold, k, o_row, o_score, o_count, o_trust, o_spk));
        SDK_CUDA(c, cudaMemsetAsync(d_flags + 1, 0, 4, c->stream));   // fallback counter consumed
    }
    return SDK_OK;
}

// An empty shard (P == 0) has nothing to score: its lists are empty, and it still joins the collective.
static int sdk_identify_any(sdk_ctx* c, const void* d_seg, const int32_t* d_seg_label, int64_t N, int32_t L, int32_t label_base,
                            int32_t pool, double threshold, int32_t k, bool allow_lazy = false) {
    if (c->P == 0) {
        SDK_TRY(sdk_reserve(c, c->goff, (size_t)(L + 1) * 8));
        SDK_TRY(sdk_launch_group_offsets(c, d_seg_label, N, L, label_base, (int64_t*)c->goff.p, c->out.flags));   // labels are still validated
        sdk_out_view v = c->out;
        v.row += (size_t)label_base * k; v.score += (size_t)label_base * k; v.spk += (size_t)label_base * k;
        v.trust += (size_t)label_base * k; v.count += label_base;
        c->last_path = 0;
        return sdk_launch_fill_empty(c, v, L, k);
    }
    return sdk_identify_core(c, d_seg, d_seg_label, N, L, label_base, pool, threshold, k, allow_lazy);
}

// Row-sharded mode: every rank MUST enter the all-gather once per identify call, or its peers block forever.  A rank
// whose local pass failed (bad labels are reported through the flags instead; this is for ENOMEM / tensor-map / CUDA
// launch errors) publishes empty lists plus a status word, joins the collective, and returns its own error afterwards;
// its peers see SDK_EPEER at fetch time.  If not even that is possible the communicator is aborted.
static int sdk_finish_collective(sdk_ctx* c, int32_t L, int32_t k, int local_rc) {
    if (c->world <= 1) return local_rc;
    const std::string local_msg = c->err;
    if (local_rc != SDK_OK) {
        bool ok = c->out_pack.p && c->out.bytes > 0 && cudaGetLastError() == cudaSuccess;
        if (ok) {
            int32_t st = local_rc;
            ok = sdk_launch_fill_empty(c, c->out, L, k) == SDK_OK &&
                 cudaMemcpyAsync(c->out.flags + SDK_FLAG_STATUS, &st, 4, cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
                 cudaStreamSynchronize(c->stream) == cudaSuccess;
        }
        if (!ok) {
            if (c->nccl_comm && g_nccl.CommAbort) { g_nccl.CommAbort(c->nccl_comm); c->nccl_comm = nullptr; }
            return sdk_fail(c, local_rc, local_msg + " (the NCCL communicator was aborted: this rank could not join the all-gather)");
        }
    }
    int r = sdk_allgather_merge(c, L, k);
    if (local_rc != SDK_OK) return sdk_fail(c, local_rc, local_msg);
    return r;
}

static int sdk_identify_dev_in(sdk_ctx* c, const void* d_seg, int32_t in_dtype, const int32_t* d_seg_label, int64_t N, int32_t L,
                               int32_t pool, double threshold, int32_t k) {
    SDK_TRY(sdk_check_identify_args(c, d_seg, d_seg_label, N, L, pool, threshold, k));
    cudaSetDevice(c->device);
    c->in_dtype = in_dtype;
    c->have_results = false;
    c->have_assign = false;
    c->lazy.active = false;                                // whatever the previous call left open is void now
    c->last_fallback = 0;
    c->last_retry = 0;
    int r = sdk_reserve_results(c, L, k);
    if (r == SDK_OK) r = sdk_identify_any(c, d_seg, d_seg_label, N, L, 0, pool, threshold, k, /*allow_lazy=*/c->world == 1);
    c->L = L; c->k = k; c->N = N;
    SDK_TRY(sdk_finish_collective(c, L, k, r));
    c->have_results = true;
    return SDK_OK;
}
int sdk_identify_dev(sdk_ctx* c, const float* d_seg, const int32_t* d_seg_label, int64_t N, int32_t L, int32_t pool,
                     double threshold, int32_t k) {
    return sdk_identify_dev_in(c, d_seg, SDK_IN_F32, d_seg_label, N, L, pool, threshold, k);
}
int sdk_identify_f16_dev(sdk_ctx* c, const uint16_t* d_seg, const int32_t* d_seg_label, int64_t N, int32_t L, int32_t pool,
                         double threshold, int32_t k) {
    return sdk_identify_dev_in(c, d_seg, SDK_IN_F16, d_seg_label, N, L, pool, threshold, k);
}

// one ncclAllGather of the head of every rank's result record, then K4 reads the gathered records in place
// (the only collective; no packing or unpacking copies)
static int sdk_allgather_merge(sdk_ctx* c, int32_t L, int32_t k) {
    if (!c->nccl_comm) return sdk_fail(c, SDK_ENCCL, "the NCCL communicator of this context was aborted");
    const size_t rec = c->out.gather_bytes;
    SDK_TRY(sdk_reserve(c, c->gather, rec * (size_t)c->world));
    sdk_prof_scope ps(c, "allgather");                   // the collective + K4 (bench.py: allgather_merge_us)
    int e = g_nccl.AllGather(c->out_pack.p, c->gather.p, rec, SDK_NCCL_CHAR, c->nccl_comm, c->stream);
    if (e != 0) return sdk_fail(c, SDK_ENCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error"));
    return sdk_launch_merge_topk(c, c->gather.p, rec, c->out, c->world, L, k);
}

// Merge of `world` per-shard result lists that were produced elsewhere (other processes / nodes, or two contexts of
// one process): host arrays in, the merged lists become this context's results (sdk_assign / sdk_results_fetch work
// on them).  Same kernel as the one that follows the all-gather.
int sdk_merge_topk(sdk_ctx* c, int32_t world, int32_t L, int32_t k, const int64_t* rows, const float* scores,
                   const int32_t* counts, const uint8_t* trust) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    if (world < 1 || world > 1024 || L < 1 || k < 1 || k > SDK_MAX_K) return sdk_fail(c, SDK_EINVAL, "merge: need 1 <= world <= 1024, L >= 1, 1 <= k <= 32");
    if (!rows || !scores || !counts) return sdk_fail(c, SDK_EINVAL, "merge: rows/scores/counts is NULL");
    cudaSetDevice(c->device);
    c->have_results = false;
    c->have_assign = false;
    c->lazy.active = false;
    SDK_TRY(sdk_reserve_results(c, L, k));
    const sdk_out_view& v = c->out;
    const size_t rec = v.gather_bytes, n = (size_t)L * k;
    std::vector<char> h(rec * (size_t)world, 0);
    for (int w = 0; w < world; ++w) {
        char* r = h.data() + rec * (size_t)w;
        memcpy(r + v.off_row, rows + (size_t)w * n, n * 8);
        memcpy(r + v.off_score, scores + (size_t)w * n, n * 4);
        memset(r + v.off_spk, 0xff, n * 4);
        memcpy(r + v.off_count, counts + (size_t)w * L, (size_t)L * 4);
        if (trust) memcpy(r + v.off_trust, trust + (size_t)w * n, n);
        else memset(r + v.off_trust, SDK_TRUST_UNKNOWN, n);
        for (int32_t g = 0; g < L; ++g) {
            const int32_t cnt = counts[(size_t)w * L + g];
            if (cnt < 0 || cnt > k) return sdk_fail(c, SDK_EINVAL, "merge: counts must be in [0,k]");
        }
    }
    SDK_TRY(sdk_reserve(c, c->gather, rec * (size_t)world));
    SDK_CUDA(c, cudaMemcpyAsync(c->gather.p, h.data(), h.size(), cudaMemcpyHostToDevice, c->stream));
    SDK_CUDA(c, cudaStreamSynchronize(c->stream));      // h goes out of scope
    SDK_TRY(sdk_launch_merge_topk(c, c->gather.p, rec, v, world, L, k));
    c->L = L; c->k = k; c->N = 0;
    c->have_results = true;
    return SDK_OK;
}

// Host-buffer entry point.  Large batches are cut at label-group boundaries into chunks that are copied on a second
// stream into a ring of three staging buffers while earlier chunks are being scored, so end-to-end time is
// max(PCIe, compute) instead of their sum, and the device footprint is three chunks instead of the whole batch.
static int sdk_identify_host_body(sdk_ctx* c, const void* seg_v, const int32_t* seg_label, int64_t N, int32_t L, int32_t pool,
                                  double threshold, int32_t k) {
    const int32_t D = c->D;
    const char* seg = (const char*)seg_v;
    const size_t row_bytes = (size_t)D * sdk_in_size(c->in_dtype);
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)((size_t)c->opt_chunk_mb << 20) / (int64_t)row_bytes);
    if (N <= chunk_rows + chunk_rows / 4) {
        SDK_TRY(sdk_reserve(c, c->seg_raw, (size_t)N * row_bytes));
        SDK_TRY(sdk_reserve(c, c->seg_lab, (size_t)N * 4));
        if (N > 0) {
            SDK_CUDA(c, cudaMemcpyAsync(c->seg_raw.p, seg, (size_t)N * row_bytes, cudaMemcpyHostToDevice, c->stream));
            SDK_CUDA(c, cudaMemcpyAsync(c->seg_lab.p, seg_label, (size_t)N * 4, cudaMemcpyHostToDevice, c->stream));
        }
        SDK_TRY(sdk_identify_any(c, c->seg_raw.p, (const int32_t*)c->seg_lab.p, N, L, 0, pool, threshold, k, /*allow_lazy=*/c->world == 1));
    } else {
        // chunk cut points: first label change at or after each multiple of chunk_rows
        for (int64_t i = 1; i < N; ++i)
            if (seg_label[i] < seg_label[i - 1]) return sdk_fail(c, SDK_EINVAL, "seg_label must be non-decreasing (segments sorted by label group)");
        if (seg_label[0] < 0 || seg_label[N - 1] >= L) return sdk_fail(c, SDK_EINVAL, "seg_label out of range [0,L)");
        std::vector<int64_t> cut{0};
        while (cut.back() < N) {
            int64_t e = std::min<int64_t>(N, cut.back() + chunk_rows);
            while (e < N && seg_label[e] == seg_label[e - 1]) ++e;
            cut.push_back(e);
        }
        int64_t max_rows = 0;
        for (size_t i = 1; i < cut.size(); ++i) max_rows = std::max(max_rows, cut[i] - cut[i - 1]);
        if (!c->copy_stream) {
            SDK_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            for (int b = 0; b < SDK_NSTAGE; ++b) {
                SDK_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
                SDK_CUDA(c, cudaEventCreateWithFlags(&c->ev_consumed[b], cudaEventDisableTiming));
            }
        }
        for (int b = 0; b < SDK_NSTAGE; ++b) {
            SDK_TRY(sdk_reserve(c, c->stage_seg[b], (size_t)max_rows * row_bytes));
            SDK_TRY(sdk_reserve(c, c->stage_lab[b], (size_t)max_rows * 4));
        }
        SDK_CUDA(c, cudaStreamSynchronize(c->stream));
        const size_t nchunk = cut.size() - 1;
        auto enqueue_copy = [&](size_t i) -> int {
            const int b = (int)(i % SDK_NSTAGE);
            const int64_t a = cut[i], n = cut[i + 1] - cut[i];
            if (i >= SDK_NSTAGE) SDK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
            SDK_CUDA(c, cudaMemcpyAsync(c->stage_seg[b].p, seg + (size_t)a * row_bytes, (size_t)n * row_bytes, cudaMemcpyHostToDevice, c->copy_stream));
            SDK_CUDA(c, cudaMemcpyAsync(c->stage_lab[b].p, seg_label + a, (size_t)n * 4, cudaMemcpyHostToDevice, c->copy_stream));
            SDK_CUDA(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
            return SDK_OK;
        };
        // three staging buffers: the copy engine runs two chunks ahead of the kernels, so a chunk whose scoring takes as long
        // as its copy (fp16 rows: twice the segments per byte) never holds the next copy back
        for (size_t i = 0; i + 1 < SDK_NSTAGE && i < nchunk; ++i) SDK_TRY(enqueue_copy(i));
        for (size_t i = 0; i < nchunk; ++i) {
            if (i + SDK_NSTAGE - 1 < nchunk) SDK_TRY(enqueue_copy(i + SDK_NSTAGE - 1));
            const int b = (int)(i % SDK_NSTAGE);
            const int64_t a = cut[i], n = cut[i + 1] - cut[i];
            const int32_t g0 = seg_label[a];
            // groups of this chunk: [g0, g1) where g1 = first label of the next chunk (empty groups in between
            // belong to this chunk); the last chunk runs to L
            const int32_t g1 = (i + 1 < nchunk) ? seg_label[cut[i + 1]] : L;
            const int32_t gbeg = (i == 0) ? 0 : g0;
            SDK_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
            SDK_TRY(sdk_identify_any(c, c->stage_seg[b].p, (const int32_t*)c->stage_lab[b].p, n, g1 - gbeg, gbeg,
                                      pool, threshold, k));
            SDK_CUDA(c, cudaEventRecord(c->ev_consumed[b], c->stream));
        }
    }
    return SDK_OK;
}

static int sdk_identify_in(sdk_ctx* c, const void* seg, int32_t in_dtype, const int32_t* seg_label, int64_t N, int32_t L, int32_t pool,
                           double threshold, int32_t k, int64_t* out_row, float* out_score, int32_t* out_count) {
    SDK_TRY(sdk_check_identify_args(c, seg, seg_label, N, L, pool, threshold, k));
    cudaSetDevice(c->device);
    c->in_dtype = in_dtype;
    c->have_results = false;
    c->have_assign = false;
    c->lazy.active = false;
    c->last_fallback = 0;
    c->last_retry = 0;
    int r = sdk_reserve_results(c, L, k);
    if (r == SDK_OK) r = sdk_identify_host_body(c, seg, seg_label, N, L, pool, threshold, k);
    c->L = L; c->k = k; c->N = N;
    SDK_TRY(sdk_finish_collective(c, L, k, r));
    c->have_results = true;
    return sdk_results_fetch(c, out_row, out_score, out_count, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}
int sdk_identify(sdk_ctx* c, const float* seg, const int32_t* seg_label, int64_t N, int32_t L, int32_t pool,
                 double threshold, int32_t k, int64_t* out_row, float* out_score, int32_t* out_count) {
    return sdk_identify_in(c, seg, SDK_IN_F32, seg_label, N, L, pool, threshold, k, out_row, out_score, out_count);
}
int sdk_identify_f16(sdk_ctx* c, const uint16_t* seg, const int32_t* seg_label, int64_t N, int32_t L, int32_t pool,
                     double threshold, int32_t k, int64_t* out_row, float* out_score, int32_t* out_count) {
    return sdk_identify_in(c, seg, SDK_IN_F16, seg_label, N, L, pool, threshold, k, out_row, out_score, out_count);
}

int sdk_assign(sdk_ctx* c, double assign_threshold, int32_t min_trust_code) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    if (!c->have_results) return sdk_fail(c, SDK_ESTATE, "sdk_assign before sdk_identify");
    cudaSetDevice(c->device);
    const int32_t L = c->L, k = c->k;
    const sdk_out_view& v = c->out;
    SDK_TRY(sdk_launch_assign(c, v.row, v.score, v.trust, v.count, L, k, assign_threshold, min_trust_code, v.as_idx,
                              v.as_score, v.as_conf, v.as_cidx, v.as_cscore));
    c->assign_thr = assign_threshold;                      // (a pending small-query certificate may make sdk_settle redo this)
    c->assign_min_trust = min_trust_code;
    c->have_assign = true;
    return SDK_OK;
}

int sdk_results_fetch(sdk_ctx* c, int64_t* out_row, float* out_score, int32_t* out_count, uint8_t* out_trust,
                      int32_t* assign_idx, double* assign_score, int32_t* assign_conf, int32_t* cand_idx,
                      double* cand_score) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    if (!c->have_results) return sdk_fail(c, SDK_ESTATE, "no results: call sdk_identify first");
    const bool want_assign = assign_idx || assign_score || assign_conf || cand_idx || cand_score;
    if (want_assign && !c->have_assign) return sdk_fail(c, SDK_ESTATE, "assignment outputs requested before sdk_assign");
    cudaSetDevice(c->device);
    const size_t n = (size_t)c->L * c->k, L = (size_t)c->L;
    const sdk_out_view& v = c->out;
    cudaStream_t s = c->stream;
    const size_t need = want_assign ? v.bytes : v.gather_bytes;
    if (need <= ((size_t)1 << 20)) {
        // small record (latency shapes): ONE device -> host copy into pinned memory, then host copies
        if (c->h_pack_cap < v.bytes) {
            if (c->h_pack) cudaFreeHost(c->h_pack);
            c->h_pack = nullptr;
            c->h_pack_cap = 0;
            const size_t cap = std::max<size_t>(v.bytes * 2, 65536);
            if (cudaHostAlloc(&c->h_pack, cap, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return sdk_fail(c, SDK_ENOMEM, "cudaHostAlloc of the result staging buffer failed");
            }
            c->h_pack_cap = cap;
        }
        SDK_CUDA(c, cudaMemcpyAsync(c->h_pack, c->out_pack.p, need, cudaMemcpyDeviceToHost, s));
        SDK_CUDA(c, cudaStreamSynchronize(s));
        if (c->lazy.active) {
            // small-query path: the record just read carries the flags; only a failed certificate costs a second copy
            const bool redo = ((const int32_t*)c->h_pack)[SDK_FLAG_FB] > 0;
            SDK_TRY(sdk_settle(c, (const int32_t*)c->h_pack));
            if (redo) {
                SDK_CUDA(c, cudaMemcpyAsync(c->h_pack, c->out_pack.p, need, cudaMemcpyDeviceToHost, s));
                SDK_CUDA(c, cudaStreamSynchronize(s));
            }
        }
        const char* h = (const char*)c->h_pack;
        SDK_TRY(sdk_flag_error(c, (const int32_t*)h));
        if (out_row) memcpy(out_row, h + v.off_row, n * 8);
        if (out_score) memcpy(out_score, h + v.off_score, n * 4);
        if (out_count) memcpy(out_count, h + v.off_count, L * 4);
        if (out_trust) memcpy(out_trust, h + v.off_trust, n);
        if (assign_idx) memcpy(assign_idx, h + v.off_as_idx, L * 4);
        if (assign_score) memcpy(assign_score, h + v.off_as_score, L * 8);
        if (assign_conf) memcpy(assign_conf, h + v.off_as_conf, L * 4);
        if (cand_idx) memcpy(cand_idx, h + v.off_as_cidx, L * 12);
        if (cand_score) memcpy(cand_score, h + v.off_as_cscore, L * 24);
        return SDK_OK;
    }
    SDK_TRY(sdk_settle(c));
    if (out_row) SDK_CUDA(c, cudaMemcpyAsync(out_row, v.row, n * 8, cudaMemcpyDeviceToHost, s));
    if (out_score) SDK_CUDA(c, cudaMemcpyAsync(out_score, v.score, n * 4, cudaMemcpyDeviceToHost, s));
    if (out_count) SDK_CUDA(c, cudaMemcpyAsync(out_count, v.count, L * 4, cudaMemcpyDeviceToHost, s));
    if (out_trust) SDK_CUDA(c, cudaMemcpyAsync(out_trust, v.trust, n, cudaMemcpyDeviceToHost, s));
    if (assign_idx) SDK_CUDA(c, cudaMemcpyAsync(assign_idx, v.as_idx, L * 4, cudaMemcpyDeviceToHost, s));
    if (assign_score) SDK_CUDA(c, cudaMemcpyAsync(assign_score, v.as_score, L * 8, cudaMemcpyDeviceToHost, s));
    if (assign_conf) SDK_CUDA(c, cudaMemcpyAsync(assign_conf, v.as_conf, L * 4, cudaMemcpyDeviceToHost, s));
    if (cand_idx) SDK_CUDA(c, cudaMemcpyAsync(cand_idx, v.as_cidx, L * 12, cudaMemcpyDeviceToHost, s));
    if (cand_score) SDK_CUDA(c, cudaMemcpyAsync(cand_score, v.as_cscore, L * 24, cudaMemcpyDeviceToHost, s));
    SDK_CUDA(c, cudaStreamSynchronize(s));
    return sdk_check_flags(c, v.flags);
}

// ---- config 5: pooled self-affinity -------------------------------------------------------------
__global__ void k_affinity_finish(const long long* __restrict__ qpool /*[L,N]*/, const int64_t* __restrict__ goff,
                                  int64_t N, int32_t L, int32_t pool, float* __restrict__ out_nl) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * L) return;
    int64_t n = i / L;
    int32_t l = (int32_t)(i - n * L);
    out_nl[i] = sdk_pool_finish(qpool[(int64_t)l * N + n], goff[l + 1] - goff[l], pool);
}
// out_ll[a,b] = mean over the segments of label a of out_nl[.,b], pooled in Q30 integers
__global__ void k_affinity_ll(const float* __restrict__ out_nl, const int64_t* __restrict__ goff, int32_t L, int64_t N,
                              float* __restrict__ out_ll) {
    const int a = blockIdx.x, b = blockIdx.y;
    int64_t s0 = goff[a], s1 = goff[a + 1];            // (clamped: on the exact path the label flag is only read at the end)
    s0 = s0 < 0 ? 0 : (s0 > N ? N : s0);
    s1 = s1 < 0 ? 0 : (s1 > N ? N : s1);
    long long acc = 0;
    for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) acc += __double2ll_rn((double)out_nl[s * L + b] * SDK_Q30);
    __shared__ long long sh[32];
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        out_ll[a * L + b] = s1 > s0 ? (float)((double)t / ((double)(s1 - s0) * SDK_Q30)) : 0.f;
    }
}

int sdk_affinity_pooled_dev(sdk_ctx* c, const float* d_seg, const int32_t* d_seg_label, int64_t N, int32_t D, int32_t L,
                            int32_t dtype, int32_t pool, float* d_out_nl, float* d_out_ll) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    if (N < 1 || D < 1 || D > 8192 || L < 1 || !d_seg || !d_seg_label || !d_out_nl)
        return sdk_fail(c, SDK_EINVAL, "affinity: bad arguments");
    if (dtype != SDK_DTYPE_F32 && dtype != SDK_DTYPE_BF16) return sdk_fail(c, SDK_EINVAL, "dtype must be 0 or 1");
    if (pool != SDK_POOL_MEAN && pool != SDK_POOL_MAX) return sdk_fail(c, SDK_EINVAL, "pool must be 0 or 1");
    cudaSetDevice(c->device);
    const bool bf16 = dtype == SDK_DTYPE_BF16;
    const int32_t Dp = (D + 63) / 64 * 64;
    SDK_TRY(sdk_reserve(c, c->goff, (size_t)(L + 1) * 8));
    SDK_TRY(sdk_reserve(c, c->flags, SDK_NFLAGS * 4));
    SDK_CUDA(c, cudaMemsetAsync(c->flags.p, 0, SDK_NFLAGS * 4, c->stream));
    SDK_TRY(sdk_launch_group_offsets(c, d_seg_label, N, L, 0, (int64_t*)c->goff.p, (int32_t*)c->flags.p));
    const double macs = (double)N * (double)N * (double)Dp;
    int path = c->opt_path;
    if (path == 0) path = (macs > 536870912.0 && bf16 && sdk_poolgemm_supported(Dp) && c->tmap_encode) ? 2 : 1;
    if (path == 2 && !(bf16 && sdk_poolgemm_supported(Dp) && c->tmap_encode))
        return sdk_fail(c, SDK_EINVAL, "tcgen05 affinity needs dtype bf16 and a supported D");
    c->last_path = path;
    if (!bf16) SDK_TRY(sdk_reserve(c, c->seg_f32, (size_t)N * D * 4));
    else SDK_TRY(sdk_reserve(c, c->seg_bf16, (size_t)N * Dp * 2));
    SDK_TRY(sdk_launch_normalize(c, d_seg, N, D, Dp, bf16 ? nullptr : (float*)c->seg_f32.p,
                                 bf16 ? (__nv_bfloat16*)c->seg_bf16.p : nullptr));
    if (path == 2) {
        SDK_CUDA(c, cudaMemsetAsync(d_out_nl, 0, (size_t)N * L * 4, c->stream));     // labels without segments: affinity 0
        // mean pooling: accumulate-pooling with the label columns split (poolacc.cu plan B); max pooling: generic kernel
        bool use_acc = c->opt_acc && sdk_poolacc_applicable(Dp, L, pool) && D % 4 == 0 && D <= 2048 && (uintptr_t)d_seg % 16 == 0;
        int64_t acc_steps = 0;
        if (use_acc) {
            int32_t lf = 0;
            SDK_CUDA(c, cudaMemcpyAsync(&lf, c->flags.p, 4, cudaMemcpyDeviceToHost, c->stream));
            SDK_CUDA(c, cudaStreamSynchronize(c->stream));
            if (lf & 1) return sdk_fail(c, SDK_EINVAL, "seg_label out of range [0,L)");
            if (lf & 2) return sdk_fail(c, SDK_EINVAL, "seg_label must be non-decreasing (segments sorted by label group)");
            SDK_TRY(sdk_poolacc_plan(c, (const int64_t*)c->goff.p, L, N, N, Dp, &acc_steps));
            if (c->opt_acc != 2 && (double)acc_steps * 256.0 > 1.12 * (double)N + 1024.0) use_acc = false;
        }
        if (use_acc) {
            c->last_path = 3;
            SDK_TRY(sdk_launch_poolacc(c, d_seg, SDK_IN_F32, d_seg_label, 0, N, D, Dp, (const __nv_bfloat16*)c->seg_bf16.p, N,
                                       (const int64_t*)c->goff.p, L, acc_steps, 1, 0.f, 0, nullptr, nullptr, d_out_nl, c->seg_il, nullptr));
        } else {
            SDK_TRY(sdk_launch_poolgemm_dense(c, (const __nv_bfloat16*)c->seg_bf16.p, N, (const __nv_bfloat16*)c->seg_bf16.p, N,
                                              Dp, (const int64_t*)c->goff.p, L, pool, d_out_nl));
        }
    } else {
        const void* ops = bf16 ? c->seg_bf16.p : c->seg_f32.p;
        SDK_TRY(sdk_reserve(c, c->qpool, (size_t)L * N * 8));
        SDK_TRY(sdk_launch_exact(c, ops, ops, bf16, D, bf16 ? Dp : D, (const int64_t*)c->goff.p, nullptr, L, nullptr, N, pool,
                                 (long long*)c->qpool.p, nullptr, N));
        sdk_prof_scope ps(c, "affinity");
        int64_t tot = N * L;
        k_affinity_finish<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>((const long long*)c->qpool.p,
                                                                                (const int64_t*)c->goff.p, N, L, pool, d_out_nl);
        c->launches++;
        SDK_CUDA(c, cudaGetLastError());
    }
    if (d_out_ll) {
        k_affinity_ll<<<dim3(L, L), 256, 0, c->stream>>>(d_out_nl, (const int64_t*)c->goff.p, L, N, d_out_ll);
        c->launches++;
        SDK_CUDA(c, cudaGetLastError());
    }
    return SDK_OK;
}

int sdk_affinity_pooled(sdk_ctx* c, const float* seg, const int32_t* seg_label, int64_t N, int32_t D, int32_t L,
                        int32_t dtype, int32_t pool, float* out_nl, float* out_ll) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    if (N < 1 || !seg || !seg_label || !out_nl) return sdk_fail(c, SDK_EINVAL, "affinity: bad arguments");
    cudaSetDevice(c->device);
    SDK_TRY(sdk_reserve(c, c->seg_raw, (size_t)N * D * 4));
    SDK_TRY(sdk_reserve(c, c->seg_lab, (size_t)N * 4));
    SDK_TRY(sdk_reserve(c, c->dense, (size_t)N * L * 4 + (size_t)L * L * 4));
    SDK_CUDA(c, cudaMemcpyAsync(c->seg_raw.p, seg, (size_t)N * D * 4, cudaMemcpyHostToDevice, c->stream));
    SDK_CUDA(c, cudaMemcpyAsync(c->seg_lab.p, seg_label, (size_t)N * 4, cudaMemcpyHostToDevice, c->stream));
    float* d_nl = (float*)c->dense.p;
    float* d_ll = out_ll ? d_nl + (size_t)N * L : nullptr;
    SDK_TRY(sdk_affinity_pooled_dev(c, (const float*)c->seg_raw.p, (const int32_t*)c->seg_lab.p, N, D, L, dtype, pool, d_nl, d_ll));
    SDK_CUDA(c, cudaMemcpyAsync(out_nl, d_nl, (size_t)N * L * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_ll) SDK_CUDA(c, cudaMemcpyAsync(out_ll, d_ll, (size_t)L * L * 4, cudaMemcpyDeviceToHost, c->stream));
    SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    return sdk_check_flags(c, (const int32_t*)c->flags.p);
}

// ---- stream / timing ----------------------------------------------------------------------------
int sdk_sync(sdk_ctx* c) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    return SDK_OK;
}
void* sdk_stream(sdk_ctx* c) { return c ? (void*)c->stream : nullptr; }
int sdk_timer_start(sdk_ctx* c) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    SDK_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
    return SDK_OK;
}
int sdk_timer_stop(sdk_ctx* c, float* ms) {
    if (!c || !ms) return sdk_fail(c, SDK_EINVAL, "NULL ctx/ms");
    SDK_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
    SDK_CUDA(c, cudaEventSynchronize(c->ev_t1));
    SDK_CUDA(c, cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return SDK_OK;
}
static void sdk_prof_drain(sdk_ctx* c) {
    if (c->pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& p : c->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            c->prof[p.name].ms += ms;
            c->prof[p.name].launches += 1;
        } else cudaGetLastError();
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    c->pending.clear();
}
int sdk_profile_get(sdk_ctx* c, const char* name, float* ms_total, int64_t* launches) {
    if (!c || !name) return sdk_fail(c, SDK_EINVAL, "NULL ctx/name");
    sdk_prof_drain(c);
    auto it = c->prof.find(name);
    if (ms_total) *ms_total = it == c->prof.end() ? 0.f : (float)it->second.ms;
    if (launches) *launches = it == c->prof.end() ? 0 : it->second.launches;
    return SDK_OK;
}
int sdk_profile_reset(sdk_ctx* c) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    sdk_prof_drain(c);
    c->prof.clear();
    return SDK_OK;
}
int64_t sdk_launch_count(sdk_ctx* c) { return c ? c->launches : 0; }
int sdk_probe_bank_read(sdk_ctx* c) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    cudaSetDevice(c->device);
    return sdk_launch_probe_read(c);
}

// Diagnostics of the certified top-k: the first-chance candidate rows of every label group with their STAGE-A
// (tensor-core / bank-stream) pooled scores, and the margin model eps_g = eps_base + eps_chain * chain_g.
int sdk_stage_a_fetch(sdk_ctx* c, int32_t* rows, float* approx, int64_t cap, int32_t* ncand, float* eps_base, float* eps_chain) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    if (!c->have_results || c->last_ncand <= 0) return sdk_fail(c, SDK_ESTATE, "no stage-A lists: the last identify did not take a tensor / bank-stream path");
    SDK_TRY(sdk_settle(c));
    const int64_t n = (int64_t)c->last_cand_groups * c->last_ncand;
    if (ncand) *ncand = c->last_ncand;
    if (eps_base) *eps_base = c->last_eps_base;
    if (eps_chain) *eps_chain = c->last_eps_chain;
    if (rows || approx) {
        if (cap < n) return sdk_fail(c, SDK_EINVAL, "stage-A fetch: buffers too small (need groups * ncand entries)");
        cudaSetDevice(c->device);
        if (rows) SDK_CUDA(c, cudaMemcpyAsync(rows, c->cand_row.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (approx) SDK_CUDA(c, cudaMemcpyAsync(approx, c->cand_val.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        SDK_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return SDK_OK;
}
int64_t sdk_last_retry(sdk_ctx* c) { return c ? c->last_retry : 0; }
int sdk_last_path(sdk_ctx* c, int32_t* path, int64_t* n_fallback) {
    if (!c) return sdk_fail(nullptr, SDK_EINVAL, "ctx is NULL");
    SDK_TRY(sdk_settle(c));                                // the fall-back count of a small-query call is known only now
    if (path) *path = c->last_path;
    if (n_fallback) *n_fallback = c->last_fallback;
    return SDK_OK;
}

}  // extern "C"
