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
