"""Seeded synthetic inputs of the shapes BASELINE.json names.

The STT / embedding extractors of the reference are network APIs, so every benchmark and parity test
runs on planted data: speaker centroid c ~ N(0, I_D) normalised; an embedding of that speaker is
normalise(c + sigma * n / sqrt(D)) times a random scale in [0.5, 20] (so the normalise kernel does real
work).  sigma = 0.35 gives same-speaker cosine ~ 0.89 and different-speaker cosine ~ 0 +- 1/sqrt(D).
A fraction of the labels are impostors (no enrolled speaker) to exercise the threshold.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np


@dataclass
class Case:
    seg: np.ndarray          # [N, D] fp32 raw segment embeddings, sorted by label group
    seg_label: np.ndarray    # [N] int32 group index, non-decreasing
    goff: np.ndarray         # [G+1] int64 CSR offsets
    bank: np.ndarray         # [P, D] fp32 raw bank rows
    row_speaker: np.ndarray  # [P] int32, contiguous per speaker
    row_trust: np.ndarray    # [P] uint8 trust codes
    truth: np.ndarray        # [G] int32 true speaker per group, -1 for impostors
    n_speakers: int

    @property
    def G(self) -> int:
        return len(self.goff) - 1


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def zipf_counts(rng, total: int, n: int) -> np.ndarray:
    """Split `total` segments over n labels with 1/rank weights (every label gets >= 1)."""
    w = 1.0 / np.arange(1, n + 1)
    rng.shuffle(w)
    c = np.maximum(1, np.floor(w / w.sum() * total).astype(np.int64))
    c[np.argmax(c)] += total - c.sum()
    return c


def make_case(seed: int, seg_counts: Sequence[int], n_speakers: int, D: int, rows_per_speaker=1, sigma: float = 0.35,
              impostor_frac: float = 0.1, neighbours: int = 0, truth: Optional[Sequence[int]] = None,
              trust_cycle: Sequence[int] = (0, 1, 2)) -> Case:
    """seg_counts[g] = number of segments of label group g.  `neighbours` plants that many extra enrolled
    speakers correlated (cos 0.5-0.8) with each true speaker's centroid (for meaningful top-k lists)."""
    rng = np.random.default_rng(seed)
    seg_counts = np.asarray(seg_counts, dtype=np.int64)
    G = len(seg_counts)
    cent = _unit(rng.standard_normal((n_speakers, D))).astype(np.float64)
    if neighbours > 0:
        for s in range(0, n_speakers - neighbours, neighbours + 1):
            for j in range(1, neighbours + 1):
                a = rng.uniform(0.5, 0.8)
                cent[s + j] = _unit(a * cent[s] + np.sqrt(1 - a * a) * cent[s + j])
    rps = np.broadcast_to(np.asarray(rows_per_speaker, dtype=np.int64), (n_speakers,)).copy()
    row_speaker = np.repeat(np.arange(n_speakers, dtype=np.int32), rps)
    P = len(row_speaker)
    bank = _unit(cent[row_speaker] + sigma * rng.standard_normal((P, D)) / np.sqrt(D))
    bank = (bank * rng.uniform(0.5, 20.0, size=(P, 1))).astype(np.float32)
    row_trust = np.asarray([trust_cycle[i % len(trust_cycle)] for i in range(P)], dtype=np.uint8)
    if truth is None:
        truth = rng.integers(0, n_speakers, size=G).astype(np.int32)
        truth[rng.random(G) < impostor_frac] = -1
    truth = np.asarray(truth, dtype=np.int32)
    goff = np.zeros(G + 1, dtype=np.int64)
    np.cumsum(seg_counts, out=goff[1:])
    N = int(goff[-1])
    seg = np.empty((N, D), dtype=np.float32)
    for g in range(G):
        n = int(seg_counts[g])
        if n == 0:
            continue
        c = cent[truth[g]] if truth[g] >= 0 else _unit(rng.standard_normal(D))
        x = _unit(c[None, :] + sigma * rng.standard_normal((n, D)) / np.sqrt(D))
        seg[goff[g]:goff[g + 1]] = (x * rng.uniform(0.5, 20.0, size=(n, 1))).astype(np.float32)
    seg_label = np.repeat(np.arange(G, dtype=np.int32), seg_counts)
    return Case(seg, seg_label, goff, bank, row_speaker, row_trust, truth, n_speakers)


# ---- the five BASELINE.json configs at (optionally reduced) size ---------------------------------
def config1(seed: int = 101) -> Case:
    """2 labels, 40 segments (S1:22, S2:18), 3 enrolled profiles (trust high/medium/low), D=192."""
    return make_case(seed, [22, 18], 3, 192, truth=[0, 1], trust_cycle=(0, 1, 2))


def config2(seed: int = 202, total: int = 2000, labels: int = 8, P: int = 500, D: int = 256) -> Case:
    rng = np.random.default_rng(seed)
    return make_case(seed, zipf_counts(rng, total, labels), P, D)


def config3(seed: int = 303, recordings: int = 16, seg_per_rec: int = 2000, labels: int = 8, P: int = 10000, D: int = 192):
    rng = np.random.default_rng(seed)
    counts = np.concatenate([zipf_counts(rng, max(labels, int(rng.poisson(seg_per_rec))), labels) for _ in range(recordings)])
    return make_case(seed, counts, P, D)


def config4(seed: int = 404, P: int = 100000, D: int = 512, labels: int = 8, total: int = 2000, neighbours: int = 11):
    rng = np.random.default_rng(seed)
    return make_case(seed, zipf_counts(rng, total, labels), P, D, neighbours=neighbours, impostor_frac=0.0)


def config5(seed: int = 505, N: int = 4096, L: int = 16, D: int = 256):
    rng = np.random.default_rng(seed)
    counts = zipf_counts(rng, N, L)
    return make_case(seed, counts, L, D, truth=list(range(L)), impostor_frac=0.0)


# ---- on-disk fixtures: a $SPEAKERS_EMBEDDINGS_DIR store + recording + transcript + sidecar ----------
def speechmatics_transcript(labels_per_segment, words_per_segment: int = 3):
    """Speechmatics v2 shaped transcript whose per-label segmentation equals `labels_per_segment`
    (words 0.35 s apart, a punctuation item, then a 0.5 s gap)."""
    results, t = [], 0.0
    for i, spk in enumerate(labels_per_segment):
        for w in range(words_per_segment):
            results.append({"type": "word", "start_time": round(t, 3), "end_time": round(t + 0.3, 3),
                            "alternatives": [{"content": f"w{i}_{w}", "confidence": 1.0, "language": "en", "speaker": spk}]})
            t += 0.35
        results.append({"type": "punctuation", "start_time": round(t - 0.05, 3), "end_time": round(t - 0.05, 3),
                        "attaches_to": "previous", "is_eos": True,
                        "alternatives": [{"content": ".", "confidence": 1.0, "language": "en", "speaker": spk}]})
        t += 0.5
    return {"format": "2.9", "metadata": {"type": "transcription"}, "results": results}


def write_store(case: Case, root, label_names, backend: str = "b200", speaker_names=None, audio_name: str = "rec.wav"):
    """Writes the profile store, the bank vectors, a fake WAV, a transcript and the segment-embedding sidecar.
    Returns (audio_path, transcript_path, speaker_ids)."""
    import json
    from pathlib import Path
    from . import store as _store
    root = Path(root)
    (root / "db").mkdir(parents=True, exist_ok=True)
    trust_names = ["high", "medium", "low", "invalidated", "unknown"]
    speaker_ids = speaker_names or [f"spk{idx:04d}" for idx in range(case.n_speakers)]
    for s, sid in enumerate(speaker_ids):
        recs = []
        for r in np.flatnonzero(case.row_speaker == s):
            emb_id = f"emb-{r:08x}"
            d = root / "embeddings" / sid
            d.mkdir(parents=True, exist_ok=True)
            np.save(d / f"{emb_id}.npy", case.bank[r])
            recs.append({"id": emb_id, "external_id": None, "model_version": f"{backend}-cosine-v1",
                         "trust_level": trust_names[int(case.row_trust[r])], "samples": {}, "created_at": "2026-01-01T00:00:00+00:00"})
        prof = {"id": sid, "version": 1, "names": {"default": sid.title()}, "nicknames": [], "description": "", "metadata": {},
                "tags": [], "embeddings": {backend: recs}, "created_at": "2026-01-01T00:00:00+00:00",
                "updated_at": "2026-01-01T00:00:00+00:00"}
        (root / "db" / f"{sid}.json").write_text(json.dumps(prof, indent=2))
    audio = root / audio_name
    audio.write_bytes(b"RIFF" + (36).to_bytes(4, "little") + b"WAVEfmt " + bytes(24) + audio_name.encode())
    # interleave the labels' segments in time so that every segment is its own transcript run
    per_label = [list(np.flatnonzero(case.seg_label == g)) for g in range(case.G)]
    order, cursor = [], [0] * case.G
    while any(cursor[g] < len(per_label[g]) for g in range(case.G)):
        for g in range(case.G):
            if cursor[g] < len(per_label[g]):
                order.append(per_label[g][cursor[g]])
                cursor[g] += 1
    seg_labels = [label_names[int(case.seg_label[i])] for i in order]
    tr = speechmatics_transcript(seg_labels)
    tpath = root / (audio_name + ".speechmatics.json")
    tpath.write_text(json.dumps(tr))
    starts = np.arange(len(order)) * (3 * 0.35 + 0.5)
    _store.save_segment_embeddings(audio, backend, case.seg[order], seg_labels, starts, starts + 0.95)
    return audio, tpath, speaker_ids
