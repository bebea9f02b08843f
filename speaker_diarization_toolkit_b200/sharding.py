"""Host-side sharding logic for the two multi-GPU modes.

  * data-parallel over recordings (config 3): `partition_recordings` -- contiguous, balanced by segment count,
    no collective on the data path;
  * bank row-sharding (config 4): `shard_bank_rows` -- cut points on SPEAKER boundaries (a speaker's rows never
    straddle two shards, so the per-rank row->speaker max is final and the merge needs no speaker de-duplication);
  * `merge_topk_lists` -- the host mirror of the device merge kernel K4 (csrc/select.cu k_merge_topk): merges
    `world` per-rank top-k lists by (-score, global row).  Used by the CPU (gloo) tests of the N>1 path and to
    cross-check the kernel; the product path merges on the device after the ncclAllGather.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def partition_recordings(seg_per_recording: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """[r0, r1) recording ranges, contiguous, segment counts as even as a contiguous split allows."""
    c = np.asarray(seg_per_recording, dtype=np.int64)
    R = len(c)
    if world < 1:
        raise ValueError("world must be >= 1")
    cum = np.concatenate([[0], np.cumsum(c)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left"))
        if i > 0 and (i > R or abs(cum[i - 1] - target) <= abs(cum[min(i, R)] - target)):
            i -= 1
        cuts.append(min(max(i, cuts[-1]), R))
    cuts.append(R)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def shard_bank_rows(row_speaker: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """[p0, p1) row ranges cut on speaker boundaries, row counts as even as those boundaries allow."""
    spk = np.asarray(row_speaker)
    P = len(spk)
    if P == 0:
        return [(0, 0)] * world
    starts = np.flatnonzero(np.r_[True, spk[1:] != spk[:-1]])          # first row of every speaker run
    bounds = np.r_[starts, P]
    if len(set(spk[starts].tolist())) != len(starts):
        raise ValueError("rows of one speaker must be contiguous in the bank")
    cuts = [0]
    for r in range(1, world):
        target = P * r / world
        i = int(np.argmin(np.abs(bounds - target)))
        cuts.append(max(int(bounds[i]), cuts[-1]))
    cuts.append(P)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def merge_topk_lists(rows, scores, counts, k: int):
    """rows [W, L, k] int64 (global rows, -1 pad), scores [W, L, k] f32, counts [W, L] -> merged
    (rows [L,k], scores [L,k], counts [L], src [L,k] = (rank, index) flattened as rank*k+index)."""
    rows, scores, counts = np.asarray(rows), np.asarray(scores), np.asarray(counts)
    W, L, kk = rows.shape
    out_r = np.full((L, k), -1, np.int64)
    out_s = np.zeros((L, k), np.float32)
    out_c = np.zeros(L, np.int32)
    out_src = np.full((L, k), -1, np.int64)
    for g in range(L):
        ent = [(-float(scores[w, g, i]), int(rows[w, g, i]), w * kk + i)
               for w in range(W) for i in range(int(counts[w, g])) if rows[w, g, i] >= 0]
        ent.sort(key=lambda e: (e[0], e[1]))
        for j, (ns, r, src) in enumerate(ent[:k]):
            out_r[g, j], out_s[g, j], out_src[g, j] = r, np.float32(-ns), src
        out_c[g] = min(k, len(ent))
    return out_r, out_s, out_c, out_src
