/*
 * sdk_b200.h -- C-ABI of the B200-native embedding-matching hot path.
 *
 * The reference (CLIAI/speaker-diarization-toolkit) is pure Python and has NO native FFI for this
 * path: the slot the path sits in is the Python plugin method
 *     EmbeddingBackend.identify_speaker(audio_path, candidates, threshold=0.354)
 *         speaker_detection_backends/base.py:130-151      (called from speaker_detection:1071)
 * plus verify_speaker (base.py:153-180) and the signal fusion that consumes its rows
 *     combine_signals            speaker-assign:418-492   (constants :49-70, min-trust filter :304-311)
 * The entry points below are what a ctypes stub inside such a backend binds (INTEGRATION.md shows
 * the stub).  Plain pointers and sizes only; no torch / numpy / CUDA types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success or a negative errno-style code; the text is available
 *     from sdk_last_error(ctx) (ctx may be NULL for errors raised before a context exists)
 *   - the caller owns every host buffer; the library owns all device memory
 *   - calls on one context are not re-entrant; one context per GPU per process
 *   - there is NO CPU fallback: without an sm_100 device sdk_create fails with SDK_ENODEV
 *   - "_dev" variants take DEVICE pointers (inputs already resident in HBM) and run
 *     asynchronously on the context's stream; the others take HOST pointers and include the
 *     host<->device copies
 *
 * Canonical arithmetic (what "score" means, bit for bit) is specified in oracle/canonical.c.
 */
#ifndef SDK_B200_H
#define SDK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDK_ABI_VERSION 3

/* error codes */
#define SDK_OK 0
#define SDK_EINVAL (-22)   /* bad argument (NULL, unsorted labels, D mismatch, k too large ...) */
#define SDK_ENOMEM (-12)   /* device or host allocation failed */
#define SDK_ENODEV (-19)   /* no sm_100 device / device index out of range */
#define SDK_ECUDA (-5)     /* CUDA runtime / driver error */
#define SDK_ENCCL (-71)    /* NCCL error */
#define SDK_ESTATE (-1)    /* call order (identify before bank_load, fetch before identify ...) */
#define SDK_EPEER (-70)    /* row-sharded mode: another rank failed in this identify call; the merged result is void */

/* operand precision of the canonical score: the bank fixes it at load time */
#define SDK_DTYPE_F32 0   /* operands = fp32 normalised vectors */
#define SDK_DTYPE_BF16 1  /* operands = normalised vectors rounded to bf16 (RNE) */

#define SDK_POOL_MEAN 0
#define SDK_POOL_MAX 1

/* trust codes, index into speaker-assign:57-63 TRUST_MULTIPLIERS */
#define SDK_TRUST_HIGH 0
#define SDK_TRUST_MEDIUM 1
#define SDK_TRUST_LOW 2
#define SDK_TRUST_INVALIDATED 3
#define SDK_TRUST_UNKNOWN 4

/* confidence band codes, speaker-assign:66-70 / :465-472 */
#define SDK_CONF_UNASSIGNED 0
#define SDK_CONF_LOW 1
#define SDK_CONF_MEDIUM 2
#define SDK_CONF_HIGH 3

#define SDK_MAX_K 32       /* largest k (matches per label) */

typedef struct sdk_ctx sdk_ctx;

/* ---- life cycle ------------------------------------------------------------------------- */
int sdk_abi_version(void);
/* number of visible sm_100 devices (0 when there is none or no driver); never computes */
int sdk_device_count(void);
/* 128-byte NCCL unique id for sdk_create(world > 1); rank 0 makes it, the host broadcasts it */
int sdk_nccl_unique_id(void* out128);
/* world == 1: nccl_uid may be NULL.  world > 1: the bank is row-sharded over `world` ranks and
 * identify ends with one ncclAllGather of the per-rank top-k + a merge. */
int sdk_create(sdk_ctx** out, int device, int world, int rank, const void* nccl_uid);
void sdk_destroy(sdk_ctx* ctx);
const char* sdk_last_error(sdk_ctx* ctx);
/* tuning / test knobs: "path" (0 auto, 1 exact SIMT, 2 tcgen05), "eps" (certificate margin),
 * "profile" (1 = record per-kernel CUDA-event times), "cand" (re-scored candidates per label),
 * "chunk_mb" (host-buffer sdk_identify: size of the H2D/compute pipeline chunks, default 128),
 * "cta_group" (0 auto | 1 | 2: tcgen05 kernel variant, 2 = CTA pairs; auto takes pairs for accumulate-pooling over many groups), "acc" (0 off | 1 auto | 2 force: pool inside the MMA accumulation),
 * "gemv" (1 = stream the bank on the CUDA cores when there are <= 8 query segments),
 * "kth" (0 off | 1 auto | 2 on: candidate flush prunes against a running per-label bound on the 64th best score),
 * "poolfirst" (0 | 1: mean pooling on the tensor path contracts the label CENTROIDS in stage A -- a different algorithm,
 *  N/L times fewer flops, HBM-bound; results are identical because stage B re-scores in the canonical arithmetic),
 * "inject_fail" (test knob: the next local identify pass returns SDK_EINVAL after its first kernels) */
int sdk_set_option(sdk_ctx* ctx, const char* key, double value);

/* ---- profile bank (replaces: the candidates list handed to identify_speaker, base.py:133,
 *      i.e. the vectors under $SPEAKERS_EMBEDDINGS_DIR/embeddings, speaker_detection:88-90) --- */
/* rows        [P,D] fp32, raw (un-normalised) enrolled vectors, row-major
 * row_speaker [P]   speaker index of each row; rows of one speaker must be contiguous
 * row_trust   [P]   SDK_TRUST_* of each row's embedding record (NULL = all unknown)
 * global_row_offset first global row id of this shard (0 when world == 1)
 * world > 1 only: P == 0 is an EMPTY shard (more ranks than speaker runs); the rank still joins every all-gather. */
int sdk_bank_load(sdk_ctx* ctx, const float* rows, const int32_t* row_speaker,
                  const uint8_t* row_trust, int64_t P, int32_t D, int32_t dtype,
                  int64_t global_row_offset);
int sdk_bank_load_dev(sdk_ctx* ctx, const float* d_rows, const int32_t* d_row_speaker,
                      const uint8_t* d_row_trust, int64_t P, int32_t D, int32_t dtype,
                      int64_t global_row_offset);

/* ---- identify (replaces identify_speaker's arithmetic, base.py:130-151) ------------------ */
/* seg        [N,D] fp32 raw per-segment embeddings
 * seg_label  [N]   label-group index of each segment in [0,L), non-decreasing.  For a batch of
 *                  recordings the group index is recording*labels_per_recording + label.
 * threshold  keep speakers whose pooled score >= threshold (speaker_detection:1501 default 0.354)
 * k          matches kept per label, 1..SDK_MAX_K, ordered by (-score, global row)
 * out_row    [L,k] global bank row of each match (the row that gave the speaker its score), -1 pad
 * out_score  [L,k] canonical pooled similarity
 * out_count  [L]   number of valid matches                                                   */
int sdk_identify(sdk_ctx* ctx, const float* seg, const int32_t* seg_label, int64_t N, int32_t L,
                 int32_t pool, double threshold, int32_t k,
                 int64_t* out_row, float* out_score, int32_t* out_count);
/* device inputs; results stay on the device until sdk_results_fetch.
 * Alignment: any 4-byte aligned float pointer is legal (a tensor view at an element offset); the 128-bit load paths
 * (vector normalise, accumulate-pooling) are taken only when d_seg is 16-byte aligned and D % 4 == 0.
 * Row-sharded mode (world > 1): every rank must make the same sequence of identify calls (same L and k); a rank whose
 * local pass fails still joins the all-gather (empty lists + a status word), returns its own error, and its peers get
 * SDK_EPEER from sdk_results_fetch.
 * Small-query calls (<= 8 segments against a big bank, the bank-stream path) return with their two kernels queued and no
 * host round trip: label errors of such a call, and the exhaustive pass for labels whose top-k certificate failed, are
 * settled by the next sdk_results_fetch / sdk_last_path / sdk_stage_a_fetch (world == 1; a row-sharded context settles
 * before its all-gather). */
int sdk_identify_dev(sdk_ctx* ctx, const float* d_seg, const int32_t* d_seg_label, int64_t N,
                     int32_t L, int32_t pool, double threshold, int32_t k);

/* Raw per-segment embeddings stored as IEEE fp16 (the sidecar's compact form; SURVEY 8b leaves the sidecar format to
 * the build).  Every element is widened to fp32 -- exactly -- before the canonical normalise, so the result equals
 * sdk_identify on the widened matrix bit for bit; host->device traffic halves.  The 64-bit load paths need d_seg
 * 8-byte aligned and D % 4 == 0 (otherwise the scalar kernel runs). */
int sdk_identify_f16(sdk_ctx* ctx, const uint16_t* seg, const int32_t* seg_label, int64_t N, int32_t L,
                     int32_t pool, double threshold, int32_t k,
                     int64_t* out_row, float* out_score, int32_t* out_count);
int sdk_identify_f16_dev(sdk_ctx* ctx, const uint16_t* d_seg, const int32_t* d_seg_label, int64_t N,
                         int32_t L, int32_t pool, double threshold, int32_t k);

/* ---- assignment (replaces combine_signals over embedding_match signals, speaker-assign:418-492,
 *      and the min-trust filter, speaker-assign:304-311), computed on the device in fp64 -------- */
/* assign_threshold: speaker-assign --threshold (default 0.3, speaker-assign:756)
 * min_trust_code:   SDK_TRUST_HIGH/MEDIUM/LOW, anything else disables the filter
 * Runs on the matches of the last identify call.  Afterwards sdk_results_fetch returns:
 *   assign_idx   [L]   index in [0,k) of the chosen match, -1 = unassigned
 *   assign_score [L]   fp64 best combined score (0.4 * trust multiplier * similarity)
 *   assign_conf  [L]   SDK_CONF_*
 *   cand_idx     [L,3] match indices of the runner-up candidates (-1 pad), cand_score [L,3] */
int sdk_assign(sdk_ctx* ctx, double assign_threshold, int32_t min_trust_code);

/* ---- merge of per-shard result lists made elsewhere (replaces nothing in the reference: it is the host-visible form
 *      of the step after the all-gather, SURVEY 8e, for shards that live in other processes or nodes) -------------- */
/* rows [world,L,k] global bank rows (-1 pad), scores [world,L,k], counts [world,L], trust [world,L,k] or NULL.
 * The merged lists, ordered by (-score, global row), become the context's results: sdk_assign / sdk_results_fetch
 * work on them.  Shards must be cut on speaker boundaries (no speaker appears in two lists). */
int sdk_merge_topk(sdk_ctx* ctx, int32_t world, int32_t L, int32_t k, const int64_t* rows,
                   const float* scores, const int32_t* counts, const uint8_t* trust);

/* copies the results of the last identify / assign to host buffers (any pointer may be NULL);
 * synchronises the stream.  out_trust [L,k] = trust code of each matched row. */
int sdk_results_fetch(sdk_ctx* ctx, int64_t* out_row, float* out_score, int32_t* out_count,
                      uint8_t* out_trust, int32_t* assign_idx, double* assign_score,
                      int32_t* assign_conf, int32_t* cand_idx, double* cand_score);

/* ---- config 5: pooled segment-segment affinity -------------------------------------------- */
/* out_nl [N,L]: pooled (mean/max over the segments of label l) cosine of segment n to label l.
 * out_ll [L,L] (may be NULL): mean over the segments of label a of out_nl[.,b].
 * Operands are bf16-rounded (dtype SDK_DTYPE_BF16) or fp32.                                   */
int sdk_affinity_pooled(sdk_ctx* ctx, const float* seg, const int32_t* seg_label, int64_t N,
                        int32_t D, int32_t L, int32_t dtype, int32_t pool, float* out_nl,
                        float* out_ll);
int sdk_affinity_pooled_dev(sdk_ctx* ctx, const float* d_seg, const int32_t* d_seg_label,
                            int64_t N, int32_t D, int32_t L, int32_t dtype, int32_t pool,
                            float* d_out_nl, float* d_out_ll);

/* ---- stream / timing helpers (bench.py times on the stream the kernels run on) ------------ */
int sdk_sync(sdk_ctx* ctx);
void* sdk_stream(sdk_ctx* ctx);                      /* cudaStream_t, for torch.cuda.ExternalStream */
int sdk_timer_start(sdk_ctx* ctx);                   /* cudaEventRecord on the context stream */
int sdk_timer_stop(sdk_ctx* ctx, float* ms);         /* records, synchronises, returns elapsed ms */
/* per-kernel accounting when option "profile" = 1: total ms and launches of kernel `name`
 * ("normalize", "poolgemm", "exact", "merge", "select", "assign", "affinity") since the last reset */
int sdk_profile_get(sdk_ctx* ctx, const char* name, float* ms_total, int64_t* launches);
int sdk_profile_reset(sdk_ctx* ctx);
/* one plain read pass over the loaded bank's bf16 operands (asynchronous, on the context stream): timed by bench.py with
 * sdk_timer_start / sdk_timer_stop as the practical ceiling of an HBM-bound pass over this bank on this GPU */
int sdk_probe_bank_read(sdk_ctx* ctx);
/* number of kernels this library launched on ctx since creation (bench.py's gpu_launches) */
int64_t sdk_launch_count(sdk_ctx* ctx);
/* which path the last identify took: 1 exact SIMT, 2 tcgen05 (pooling in the epilogue), 3 tcgen05 accumulate-pooling
 * (mean pooling inside the MMA accumulation), 4 bank-stream GEMV (<= 8 query segments), 5 pool-first (centroid stage A);
 * *n_fallback = label groups whose
 * top-k certificate failed and were re-done exhaustively */
int sdk_last_path(sdk_ctx* ctx, int32_t* path, int64_t* n_fallback);
/* Diagnostics of the certified top-k (tests, bench.py): after an identify that took path 2-4 on ONE chunk, the
 * first-chance candidate rows [L, *ncand] (shard-local row index, -1 pad) with their stage-A pooled scores, and the
 * certificate's margin model: a row left out of the list has a canonical score <= its stage-A score + eps_base +
 * eps_chain * chain (chain = segments per accumulator column, or n/32 + 70 with epilogue mean pooling).  cap = entries
 * the two buffers hold (rows / approx may be NULL to read only the model). */
int sdk_stage_a_fetch(sdk_ctx* ctx, int32_t* rows, float* approx, int64_t cap, int32_t* ncand,
                      float* eps_base, float* eps_chain);
/* label groups of the last identify whose top-k certificate failed on the first candidate list and was settled by the
 * second, wider one (64 candidates) without an exhaustive pass */
int64_t sdk_last_retry(sdk_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SDK_B200_H */
