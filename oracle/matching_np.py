"""NumPy fp32 restatement of the matching path (TEST INFRASTRUCTURE / CPU baseline, never product).

Follows SURVEY.md section 8(c) steps 1-6.  The reference has no numeric matching code (its only
backend is a network API, speaker_detection_backends/speechmatics_backend.py:361-489), so steps 1-5
are "parity unpinned"; step 6 restates the reference's own `combine_signals`
(speaker-assign:418-492) and is pinned by tests/golden/combine_signals_golden.json.

This is the "what a NumPy user of the toolkit would write" version: OpenBLAS sgemm with all host
threads.  bench.py times it as the cpu_baseline ("port") and as `--impl reference`.
oracle/canonical.c is the bit-exact checker; this file agrees with it to ~1e-7 (tests/test_oracle.py).
"""
from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

# speaker-assign:49-70
SIGNAL_WEIGHTS = {"embedding_match": 0.4, "llm_name_detection": 0.3, "context_expected": 0.2,
                  "cross_backend_agreement": 0.1}
TRUST_MULTIPLIERS = {"high": 1.0, "medium": 0.7, "low": 0.4, "invalidated": 0.0, "unknown": 0.5}
CONFIDENCE_THRESHOLDS = {"high": 0.7, "medium": 0.4, "low": 0.2}
TRUST_CODES = ["high", "medium", "low", "invalidated", "unknown"]


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (nearest even) and back, NaN-free inputs."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    lsb = (u >> np.uint32(16)) & np.uint32(1)
    r = (u + np.uint32(0x7FFF) + lsb) & np.uint32(0xFFFF0000)
    return r.view(np.float32)


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """Step 1: x / max(||x||, eps), fp32 (np.linalg.norm on fp32)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = np.linalg.norm(x, axis=1).astype(np.float32)
    return x / np.maximum(n, np.float32(eps))[:, None]


def pooled_scores(seg_raw, goff, bank_raw, mode=0, pool=0, chunk_pairs=1 << 28) -> np.ndarray:
    """Steps 1-3: [G,P] pooled cosine.  mode 1 rounds the normalised operands to bf16 first.
    One sgemm per chunk of whole label groups (S stays under chunk_pairs floats), then a segmented
    sum / max over the rows of each group."""
    xs, bs = l2_normalize(seg_raw), l2_normalize(bank_raw)
    if mode == 1:
        xs, bs = bf16_round(xs), bf16_round(bs)
    goff = np.asarray(goff, dtype=np.int64)
    G, P = len(goff) - 1, bs.shape[0]
    out = np.zeros((G, P), dtype=np.float32)
    bt = np.ascontiguousarray(bs.T)
    max_rows = max(1, chunk_pairs // max(P, 1))
    g = 0
    while g < G:
        h = g + 1                                         # groups [g, h): as many whole groups as fit
        while h < G and goff[h + 1] - goff[g] <= max_rows:
            h += 1
        s0, s1 = int(goff[g]), int(goff[h])
        if s1 - s0 > max_rows:                            # one oversized group: stream it
            acc = np.zeros(P, np.float32) if pool == 0 else np.full(P, -np.inf, np.float32)
            for a in range(s0, s1, max_rows):
                blk = xs[a:min(a + max_rows, s1)] @ bt
                acc = acc + blk.sum(axis=0, dtype=np.float32) if pool == 0 else np.maximum(acc, blk.max(axis=0))
            out[g] = acc / np.float32(s1 - s0) if pool == 0 else acc
        elif s1 > s0:
            S = xs[s0:s1] @ bt
            cnt = np.diff(goff[g:h + 1])
            nz = np.flatnonzero(cnt > 0)
            starts = (goff[g:h][nz] - s0).astype(np.int64)
            if pool == 0:                                 # segmented mean as a second (tiny) sgemm
                M = np.zeros((len(nz), s1 - s0), dtype=np.float32)
                for i, j in enumerate(nz):
                    M[i, starts[i]:starts[i] + cnt[j]] = np.float32(1.0) / np.float32(cnt[j])
                out[g + nz] = M @ S
            else:
                out[g + nz] = np.maximum.reduceat(S, starts, axis=0)
        g = h
    return out


def select_topk(sim, gcount, row_speaker, threshold=0.354, k=10, row_offset=0):
    """Steps 4-5: row->speaker max (ties: lowest row), keep >= threshold, order (-score,row), top-k."""
    G, P = sim.shape
    row_speaker = np.asarray(row_speaker)
    out_row = np.full((G, k), -1, dtype=np.int64)
    out_score = np.zeros((G, k), dtype=np.float32)
    out_count = np.zeros(G, dtype=np.int32)
    order = np.lexsort((np.arange(P), row_speaker))  # by speaker then row
    spk_sorted = row_speaker[order]
    starts = np.flatnonzero(np.r_[True, spk_sorted[1:] != spk_sorted[:-1]])
    runlen = np.diff(np.r_[starts, P])
    pos = np.arange(P)
    for g in range(G):
        if gcount[g] <= 0:
            continue
        v = sim[g][order]
        best = np.maximum.reduceat(v, starts)
        # arg row: first (lowest) row attaining the max inside each speaker's run
        first = np.minimum.reduceat(np.where(v == np.repeat(best, runlen), pos, P), starts)
        arg = order[first]
        keep = best.astype(np.float64) >= threshold
        best, arg = best[keep], arg[keep]
        idx = np.lexsort((arg, -best))[:k]
        out_row[g, :len(idx)] = arg[idx] + row_offset
        out_score[g, :len(idx)] = best[idx]
        out_count[g] = len(idx)
    return out_row, out_score, out_count


def identify(seg_raw, goff, bank_raw, row_speaker, mode=0, pool=0, threshold=0.354, k=10, row_offset=0):
    sim = pooled_scores(seg_raw, goff, bank_raw, mode, pool)
    gcount = np.diff(np.asarray(goff))
    return select_topk(sim, gcount, row_speaker, threshold, k, row_offset)


# ---------------------------------------------------------------------------------------------
# Step 6: restatement of the reference's signal fusion (speaker-assign:249-258, :304-311, :418-492)
# ---------------------------------------------------------------------------------------------
@dataclass
class Signal:
    type: str
    speaker_id: Optional[str]
    score: float
    evidence: dict = field(default_factory=dict)


def passes_min_trust(trust: str, min_trust: str) -> bool:
    """speaker-assign:304-311 -- only low/medium/high take part in the ordering."""
    order = ["low", "medium", "high"]
    if min_trust in order and trust in order:
        return order.index(trust) >= order.index(min_trust)
    return True


def combine_signals(label: str, signals, threshold: float = 0.5) -> dict:
    """speaker-assign:418-492.  Returns a plain dict with the Assignment fields."""
    scores = defaultdict(float)
    evidence = defaultdict(list)
    for s in signals:
        if s.speaker_id is None:
            continue
        w = SIGNAL_WEIGHTS.get(s.type, 0.1)
        if s.type == "embedding_match":
            w *= TRUST_MULTIPLIERS.get(s.evidence.get("trust_level", "unknown"), 0.5)
        scores[s.speaker_id] += w * s.score
        evidence[s.speaker_id].append({"type": s.type, "score": s.score, **s.evidence})
    if not scores:
        return dict(speaker_label=label, speaker_id=None, confidence="unassigned", score=0.0, signals=[],
                    candidates=[])
    ranked = sorted(scores.items(), key=lambda kv: kv[1], reverse=True)
    best_id, best = ranked[0]
    if best >= CONFIDENCE_THRESHOLDS["high"]:
        conf = "high"
    elif best >= CONFIDENCE_THRESHOLDS["medium"]:
        conf = "medium"
    elif best >= CONFIDENCE_THRESHOLDS["low"]:
        conf = "low"
    else:
        conf = "unassigned"
    if best < threshold:
        return dict(speaker_label=label, speaker_id=None, confidence="unassigned", score=best,
                    signals=evidence.get(best_id, []),
                    candidates=[{"speaker_id": i, "score": v} for i, v in ranked[:3]])
    return dict(speaker_label=label, speaker_id=best_id, confidence=conf, score=best,
                signals=evidence.get(best_id, []),
                candidates=[{"speaker_id": i, "score": v} for i, v in ranked[1:4]])
