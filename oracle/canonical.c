/*
 * oracle/canonical.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C; scalar per pair, bank rows spread over the host threads) of the embedding-matching path that
 * BASELINE.json's north_star names: L2-normalise -> cosine vs profile bank -> pool per
 * diarization label -> row->speaker max -> threshold / top-k -> label->profile assignment.
 *
 * PARITY STATUS: the reference (CLIAI/speaker-diarization-toolkit) ships NO numeric matching
 * code -- its only backend delegates to the Speechmatics HTTP API
 * (speaker_detection_backends/speechmatics_backend.py:361-489, constant confidence at :486).
 * Steps 1-5 below are therefore "parity unpinned" by the reference: they follow the
 * restatement that SURVEY.md section 8(c) defines.  Step 6 (assignment) IS pinned: it restates
 * speaker-assign:418-492 (combine_signals) + :49-70 (constants) + :304-311 (min-trust filter)
 * and is checked against golden vectors produced by the reference's own combine_signals
 * (tests/golden/combine_signals_golden.json, generator tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path never does.
 *
 * "Canonical arithmetic".  Tensor cores and NumPy/OpenBLAS accumulate in different orders, so
 * scores differ in the last bits and near-ties could flip an arg-max.  The contract
 * (BASELINE.json) wants ids bit-exact.  The oracle therefore fixes ONE arithmetic that a CPU and
 * a GPU can both reproduce bit for bit, and the CUDA path re-scores its candidates in exactly
 * this arithmetic before it orders them:
 *   (1) squared norm: fp64, 32 interleaved partial sums (element e goes to partial (e/4)%32,
 *       ascending e), combined by a butterfly (xor 16,8,4,2,1); norm = (float)sqrt(.);
 *       x_hat[e] = x[e] * (1.0f / max(norm, 1e-12f)), both IEEE fp32 (one reciprocal per row, one
 *       rounding per element; within 1 ulp of NumPy's x / norm).      [SURVEY 8(c) step 1]
 *   (2) operand rounding: mode 0 keeps x_hat in fp32; mode 1 rounds it to bf16 (RNE).
 *                                                                     [SURVEY 8(c) step 2]
 *   (3) pair score: fp64 fma chain over ascending d (products of fp32/bf16 operands are exact
 *       in fp64), then fixed point q = llrint(score * 2^30).  Integer pooling makes the
 *       per-label sum/max independent of summation order.
 *   (4) pooling per label: mean -> (float)((double)sum_q / (n * 2^30)); max -> (float)(max_q / 2^30).
 *                                                                     [SURVEY 8(c) step 3]
 *   (5) speaker score = max over the speaker's bank rows (ties: lowest row); keep
 *       (double)score >= threshold; order by (-score, row); take k.   [SURVEY 8(c) steps 4-5]
 *   (6) assignment: fp64, exactly Python's float arithmetic of combine_signals.
 * |canonical - NumPy fp32| is ~1e-7 (checked in tests/test_oracle.py with tolerance 1e-5 rel).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_Q30 1073741824.0

/* ---- bf16 helpers (round to nearest even; NaN -> 0x7fff like __float2bfloat16_rn) ---- */
static inline uint16_t orc_f32_to_bf16_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}
static inline float orc_bf16_bits_to_f32(uint16_t b) {
    uint32_t u = ((uint32_t)b) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
float orc_bf16_round(float f) { return orc_bf16_bits_to_f32(orc_f32_to_bf16_bits(f)); }

/* ---- step 1+2: canonical L2 normalisation (zero rows stay zero) ---- */
void orc_normalize(const float* x, int64_t n, int32_t D, int32_t mode,
                   float* out /*[n,D]*/, float* norms /*[n] or NULL*/) {
    for (int64_t r = 0; r < n; ++r) {
        const float* xr = x + r * (int64_t)D;
        double part[32];
        for (int j = 0; j < 32; ++j) part[j] = 0.0;
        for (int32_t e = 0; e < D; ++e) {
            double v = (double)xr[e];
            int lane = (e >> 2) & 31;
            part[lane] = fma(v, v, part[lane]);
        }
        for (int off = 16; off >= 1; off >>= 1) {
            double tmp[32];
            for (int j = 0; j < 32; ++j) tmp[j] = part[j] + part[j ^ off];
            memcpy(part, tmp, sizeof(part));
        }
        float nrm = (float)sqrt(part[0]);
        if (norms) norms[r] = nrm;
        float den = nrm > 1e-12f ? nrm : 1e-12f;
        float inv = 1.0f / den;
        for (int32_t e = 0; e < D; ++e) {
            float v = xr[e] * inv;
            out[r * (int64_t)D + e] = mode == 1 ? orc_bf16_round(v) : v;
        }
    }
}

/* ---- step 3: one pair, fixed-point Q2.30 of the fp64 sequential dot ---- */
int64_t orc_pair_q30(const float* a, const float* b, int32_t D) {
    double acc = 0.0;
    for (int32_t d = 0; d < D; ++d) acc = fma((double)a[d], (double)b[d], acc);
    return (int64_t)llrint(acc * ORC_Q30);
}

static inline float orc_pool_finish(int64_t q, int64_t n, int32_t pool) {
    if (n <= 0) return 0.0f;
    if (pool == 0) return (float)((double)q / ((double)n * ORC_Q30));
    return (float)((double)q / ORC_Q30);
}

/* ---- steps 3+4: pooled canonical similarity for every (label group, bank row) ----
 * seg  [N,D]  operands (already normalised+rounded), sorted by group
 * goff [G+1]  CSR offsets of the groups into seg
 * bank [P,D]  operands
 * out  [G,P]  fp32 pooled similarity (0 for an empty group)
 * pool: 0 mean, 1 max */
/* Every (group, row) cell is independent and every pair keeps its own sequential fma chain, so the
 * order in which cells are visited cannot change a bit of the result.  The code below only arranges
 * the work for speed (bench.py checks sampled label groups of the full-size workloads against it):
 * blocks of ORC_PB bank rows are dealt to ORC_THREADS host threads (pthreads; ORC_THREADS from the
 * environment, default = online cores, 1 = the plain loop), and the ORC_PB rows of a block are scored
 * side by side against one segment so that ORC_PB independent chains fill the FMA pipeline. */
enum { ORC_PB = 8 };
typedef struct {
    const float* seg; const float* bank; float* out;
    int64_t s0, s1, P, blk0, blk1;
    int32_t D, pool;
} orc_pool_job;
static void* orc_pool_worker(void* arg) {
    const orc_pool_job* w = (const orc_pool_job*)arg;
    const int32_t D = w->D, pool = w->pool;
    const int64_t n = w->s1 - w->s0;
    double* bt = (double*)malloc(sizeof(double) * (size_t)(D > 0 ? D : 1) * ORC_PB);
    for (int64_t blk = w->blk0; blk < w->blk1; ++blk) {
        int64_t p0 = blk * ORC_PB;
        int nb = (int)(w->P - p0 < ORC_PB ? w->P - p0 : ORC_PB);
        int64_t acc[ORC_PB];
        /* the block's rows, widened to fp64 and transposed to [e][row] so that the compiler can keep the
         * ORC_PB chains in SIMD lanes (each lane is still one pair's own ascending-d fma chain) */
        for (int j = 0; j < ORC_PB; ++j) {
            acc[j] = (pool == 0) ? 0 : INT64_MIN;
            const float* bj = w->bank + (p0 + (j < nb ? j : 0)) * (int64_t)D;
            for (int32_t e = 0; e < D; ++e) bt[(int64_t)e * ORC_PB + j] = (double)bj[e];
        }
        for (int64_t s = w->s0; s < w->s1; ++s) {
            const float* a = w->seg + s * (int64_t)D;
            double d[ORC_PB];
            for (int j = 0; j < ORC_PB; ++j) d[j] = 0.0;
            for (int32_t e = 0; e < D; ++e) {
                const double ae = (double)a[e];
                const double* be = bt + (int64_t)e * ORC_PB;
                for (int j = 0; j < ORC_PB; ++j) d[j] = fma(ae, be[j], d[j]);
            }
            for (int j = 0; j < ORC_PB; ++j) {
                int64_t q = (int64_t)llrint(d[j] * ORC_Q30);     /* == orc_pair_q30(a, bank row, D) */
                if (pool == 0) acc[j] += q; else if (q > acc[j]) acc[j] = q;
            }
        }
        for (int j = 0; j < nb; ++j) w->out[p0 + j] = orc_pool_finish(acc[j], n, pool);
    }
    free(bt);
    return NULL;
}
static int orc_threads(void) {
    const char* e = getenv("ORC_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}
void orc_pooled(const float* seg, const int64_t* goff, int32_t G, const float* bank, int64_t P,
                int32_t D, int32_t pool, float* out) {
    const int nt_max = orc_threads();
    for (int32_t g = 0; g < G; ++g) {
        const int64_t nblk = (P + ORC_PB - 1) / ORC_PB;
        int nt = nt_max;
        /* small cells are not worth a thread each */
        if ((double)(goff[g + 1] - goff[g]) * (double)P * (double)D < 4.0e6) nt = 1;
        if (nt > nblk) nt = (int)(nblk > 0 ? nblk : 1);
        orc_pool_job jobs[256];
        pthread_t th[256];
        for (int t = 0; t < nt; ++t) {
            orc_pool_job* w = &jobs[t];
            w->seg = seg; w->bank = bank; w->out = out + (int64_t)g * P;
            w->s0 = goff[g]; w->s1 = goff[g + 1]; w->P = P; w->D = D; w->pool = pool;
            w->blk0 = nblk * t / nt; w->blk1 = nblk * (t + 1) / nt;
        }
        int started = 0;
        for (int t = 1; t < nt; ++t) {
            if (pthread_create(&th[t], NULL, orc_pool_worker, &jobs[t]) != 0) break;
            started = t;
        }
        orc_pool_worker(&jobs[0]);
        for (int t = 1; t <= started; ++t) pthread_join(th[t], NULL);
        for (int t = started + 1; t < nt; ++t) orc_pool_worker(&jobs[t]);   /* thread creation failed: do it here */
    }
}

/* ---- step 5: row->speaker max, threshold, order by (-score,row), top-k ----
 * sim          [G,P] pooled similarity
 * gcount       [G]   segments per group (an empty group yields no matches)
 * row_speaker  [P]   speaker index in [0,S)
 * out_row/out_score [G,k], out_count [G]; unused slots: row -1, score 0 */
typedef struct { float s; int64_t row; } orc_cand;
static int orc_cand_cmp(const void* a, const void* b) {
    const orc_cand* x = (const orc_cand*)a; const orc_cand* y = (const orc_cand*)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->row > y->row) - (x->row < y->row);
}
void orc_select(const float* sim, const int64_t* gcount, int32_t G, int64_t P,
                const int32_t* row_speaker, int32_t S, double threshold, int32_t k,
                int64_t row_offset, int64_t* out_row, float* out_score, int32_t* out_count) {
    float* best = (float*)malloc(sizeof(float) * (size_t)S);
    int64_t* arg = (int64_t*)malloc(sizeof(int64_t) * (size_t)S);
    orc_cand* c = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)S);
    for (int32_t g = 0; g < G; ++g) {
        for (int32_t i = 0; i < k; ++i) { out_row[(int64_t)g * k + i] = -1; out_score[(int64_t)g * k + i] = 0.0f; }
        out_count[g] = 0;
        if (gcount[g] <= 0) continue;
        for (int32_t s = 0; s < S; ++s) { arg[s] = -1; best[s] = 0.0f; }
        for (int64_t p = 0; p < P; ++p) {
            int32_t s = row_speaker[p];
            float v = sim[(int64_t)g * P + p];
            if (arg[s] < 0 || v > best[s]) { best[s] = v; arg[s] = p; }
        }
        int32_t m = 0;
        for (int32_t s = 0; s < S; ++s)
            if (arg[s] >= 0 && (double)best[s] >= threshold) { c[m].s = best[s]; c[m].row = arg[s]; ++m; }
        qsort(c, (size_t)m, sizeof(orc_cand), orc_cand_cmp);
        int32_t take = m < k ? m : k;
        for (int32_t i = 0; i < take; ++i) {
            out_row[(int64_t)g * k + i] = c[i].row + row_offset;
            out_score[(int64_t)g * k + i] = c[i].s;
        }
        out_count[g] = take;
    }
    free(best); free(arg); free(c);
}

/* ---- step 6: assignment = combine_signals over embedding_match signals only ----
 * Restates speaker-assign:418-492 for a signal list that holds one "embedding_match" signal
 * per returned speaker, in the order identify returned them (descending similarity):
 *   weight = SIGNAL_WEIGHTS["embedding_match"] (0.4, :49-54) * TRUST_MULTIPLIERS[trust] (:57-63)
 *   scores[id] += weight * score                                  (:431-443)
 *   stable sort descending (ties keep first inserted, :461)
 *   bands >=0.7 high / >=0.4 medium / >=0.2 low / else unassigned (:465-472)
 *   best < threshold -> id None, "unassigned", candidates = top-3 incl. best (:475-483)
 *   else candidates = ranks 2..4                                  (:485-492)
 * and the min-trust filter of collect_embedding_signals (:304-311): a signal is dropped iff
 * min_trust and its trust are both in [low, medium, high] and trust ranks lower.
 * trust codes: 0 high, 1 medium, 2 low, 3 invalidated, 4 unknown.
 * min_trust_code: 0 high, 1 medium, 2 low, anything else = filter disabled.
 * conf codes out: 0 unassigned, 1 low, 2 medium, 3 high.
 * Inputs per group g: match_row[g,k] (index into the match list, -1 = none), match_score,
 * match_trust[g,k].  Outputs: assign_idx[g] = index i in [0,k) of the chosen match or -1,
 * assign_score[g] (double), assign_conf[g], cand_idx[g,3] (-1 padded), cand_score[g,3]. */
static const double ORC_TRUST_MULT[5] = {1.0, 0.7, 0.4, 0.0, 0.5};
static int orc_trust_rank(int code) { /* index in trust_order [low, medium, high]; -1 if absent */
    if (code == 2) return 0;
    if (code == 1) return 1;
    if (code == 0) return 2;
    return -1;
}
void orc_assign(const int64_t* match_row, const float* match_score, const uint8_t* match_trust,
                const int32_t* match_count, int32_t G, int32_t k, double assign_threshold,
                int32_t min_trust_code, int32_t* assign_idx, double* assign_score,
                int32_t* assign_conf, int32_t* cand_idx, double* cand_score) {
    int min_rank = orc_trust_rank(min_trust_code);
    double* w = (double*)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1));
    int32_t* id = (int32_t*)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
    for (int32_t g = 0; g < G; ++g) {
        int32_t m = 0;
        for (int32_t i = 0; i < match_count[g]; ++i) {
            if (match_row[(int64_t)g * k + i] < 0) continue;
            int t = match_trust[(int64_t)g * k + i];
            int tr = orc_trust_rank(t);
            if (min_rank >= 0 && tr >= 0 && tr < min_rank) continue;
            double weight = 0.4;
            weight *= ORC_TRUST_MULT[t > 4 ? 4 : t];
            double ws = weight * (double)match_score[(int64_t)g * k + i];
            w[m] = 0.0 + ws;
            id[m] = i;
            ++m;
        }
        for (int j = 0; j < 3; ++j) { cand_idx[g * 3 + j] = -1; cand_score[g * 3 + j] = 0.0; }
        if (m == 0) { assign_idx[g] = -1; assign_score[g] = 0.0; assign_conf[g] = 0; continue; }
        /* stable insertion sort, descending */
        for (int32_t a = 1; a < m; ++a) {
            double wa = w[a]; int32_t ia = id[a]; int32_t b = a - 1;
            while (b >= 0 && w[b] < wa) { w[b + 1] = w[b]; id[b + 1] = id[b]; --b; }
            w[b + 1] = wa; id[b + 1] = ia;
        }
        double best = w[0];
        int conf = best >= 0.7 ? 3 : best >= 0.4 ? 2 : best >= 0.2 ? 1 : 0;
        assign_score[g] = best;
        if (best < assign_threshold) {
            assign_idx[g] = -1; assign_conf[g] = 0;
            for (int j = 0; j < 3 && j < m; ++j) { cand_idx[g * 3 + j] = id[j]; cand_score[g * 3 + j] = w[j]; }
        } else {
            assign_idx[g] = id[0]; assign_conf[g] = conf;
            for (int j = 0; j < 3 && j + 1 < m; ++j) { cand_idx[g * 3 + j] = id[j + 1]; cand_score[g * 3 + j] = w[j + 1]; }
        }
    }
    free(w); free(id);
}

/* ---- whole path in one call (what the parity tests compare the C-ABI against) ---- */
void orc_identify(const float* seg_raw, const int64_t* goff, int32_t G, int64_t N,
                  const float* bank_raw, const int32_t* row_speaker, int32_t S, int64_t P,
                  int32_t D, int32_t mode, int32_t pool, double threshold, int32_t k,
                  int64_t row_offset, int64_t* out_row, float* out_score, int32_t* out_count) {
    float* seg = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1) * (size_t)D);
    float* bank = (float*)malloc(sizeof(float) * (size_t)(P > 0 ? P : 1) * (size_t)D);
    float* sim = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1) * (size_t)(P > 0 ? P : 1));
    int64_t* gc = (int64_t*)malloc(sizeof(int64_t) * (size_t)(G > 0 ? G : 1));
    orc_normalize(seg_raw, N, D, mode, seg, NULL);
    orc_normalize(bank_raw, P, D, mode, bank, NULL);
    orc_pooled(seg, goff, G, bank, P, D, pool, sim);
    for (int32_t g = 0; g < G; ++g) gc[g] = goff[g + 1] - goff[g];
    orc_select(sim, gc, G, P, row_speaker, S, threshold, k, row_offset, out_row, out_score, out_count);
    free(seg); free(bank); free(sim); free(gc);
}

/* ---- config 5: pooled self-affinity, out[N,L]: affinity of every segment to every label ---- */
void orc_affinity(const float* seg_raw, const int64_t* goff, int32_t G, int64_t N, int32_t D,
                  int32_t mode, int32_t pool, float* out_nl /*[N,G]*/) {
    float* seg = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1) * (size_t)D);
    float* sim = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1) * (size_t)(N > 0 ? N : 1));
    orc_normalize(seg_raw, N, D, mode, seg, NULL);
    orc_pooled(seg, goff, G, seg, N, D, pool, sim); /* sim[g, row] */
    for (int32_t g = 0; g < G; ++g)
        for (int64_t r = 0; r < N; ++r) out_nl[r * (int64_t)G + g] = sim[(int64_t)g * N + r];
    free(seg); free(sim);
}
