"""Batch identify / assign over many recordings with ONE resident backend call.

The reference fans recordings out over a thread pool and runs a process tree per recording
(`speaker-process:627-642` -> `speaker-assign assign ... --use-embeddings` -> one `speaker_detection identify`
subprocess per label, `speaker-assign:283-294`).  Here the per-segment embeddings of all recordings are concatenated
(label group = recording * labels + label, which is exactly the `seg_label` convention of `sdk_identify`), scored
against the enrolled bank in one pipelined C-ABI call, and -- when embeddings are the only signal -- assigned on the
device (`sdk_assign`, the fp64 restatement of `combine_signals`).  Per recording the output is the same `mappings`
dict / assignments YAML that `speaker-assign assign` writes.

Multi-GPU, one process per GPU, both reachable from `speaker-assign assign-batch --gpus N`:
  * data-parallel over recordings (default): the manifest is cut with `sharding.partition_recordings`, every rank
    holds the whole bank and scores its own recordings -- no collective;
  * row-sharded bank (`--shard-bank`, million-profile banks): every rank loads its slice of the bank rows (cut on
    speaker boundaries, `sharding.shard_bank_rows`) and scores ALL recordings against it; the library's identify ends
    with the one ncclAllGather + merge, so every rank ends up with the global top-k (rank 0 writes the output).
`SPEAKER_B200_WORLD` / `SPEAKER_B200_RANK` / `SPEAKER_B200_UID_FILE` select the row-sharded mode for any process that
builds a `BatchMatcher` or the plugin `Backend` (rank 0 writes the NCCL unique id to the file, the others wait for it).
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import BACKEND_NAME, _native, signals as sg, store


@dataclass
class RecordingResult:
    audio_path: Path
    labels: List[str]
    matches: Dict[str, List[Dict[str, Any]]]      # label -> identify rows (descending score)
    mappings: Dict[str, Dict[str, Any]]           # label -> the `mappings` entry of speaker-assign


def sharded_context_from_env(device: int):
    """(world, rank, Context) for the row-sharded mode described by SPEAKER_B200_WORLD / _RANK / _UID_FILE, or
    (1, 0, Context(device)).  Rank 0 creates the NCCL unique id and publishes it through the file (written to a
    temporary name first, then renamed, so a reader never sees a partial id)."""
    world = int(os.environ.get("SPEAKER_B200_WORLD", "1"))
    rank = int(os.environ.get("SPEAKER_B200_RANK", "0"))
    if world <= 1:
        return 1, 0, _native.Context(device)
    uid_file = os.environ.get("SPEAKER_B200_UID_FILE")
    if not uid_file:
        raise ValueError("SPEAKER_B200_WORLD > 1 needs SPEAKER_B200_UID_FILE (a path every rank can read)")
    path = Path(uid_file)
    if rank == 0:
        uid = _native.nccl_unique_id()
        tmp = path.with_name(path.name + f".{os.getpid()}.tmp")
        tmp.write_bytes(uid)
        os.replace(tmp, path)
    else:
        deadline = time.time() + float(os.environ.get("SPEAKER_B200_UID_TIMEOUT", "120"))
        while not (path.exists() and path.stat().st_size == 128):
            if time.time() > deadline:
                raise TimeoutError(f"rank {rank}: no NCCL unique id in {path}")
            time.sleep(0.02)
        uid = path.read_bytes()
    return world, rank, _native.Context(device, world, rank, uid)


class BatchMatcher:
    """Keeps the device context and the loaded bank across calls.  `world` > 1 (with `nccl_uid`) = one rank of a
    row-sharded bank; without arguments the SPEAKER_B200_WORLD / _RANK / _UID_FILE environment decides."""

    def __init__(self, backend_name: str = BACKEND_NAME, device: Optional[int] = None, dtype: Optional[str] = None,
                 pool: Optional[str] = None, k: Optional[int] = None, threshold: float = 0.354, world: Optional[int] = None,
                 rank: int = 0, nccl_uid: Optional[bytes] = None):
        self.backend_name = backend_name
        dev = int(os.environ.get("SPEAKER_B200_DEVICE", 0)) if device is None else device
        if world is None:
            self.world, self.rank, self.ctx = sharded_context_from_env(dev)
        else:
            self.world, self.rank = int(world), int(rank)
            self.ctx = _native.Context(dev, self.world, self.rank, nccl_uid) if self.world > 1 else _native.Context(dev)
        if os.environ.get("SPEAKER_B200_STAGE_A", "") == "poolfirst":
            # mean pooling: stage A contracts the label centroids instead of the segments (a different algorithm with the
            # same, certified result: DESIGN.md section 5); several times less time to solution on long recordings
            self.ctx.set_option("poolfirst", 1)
        self.dtype = _native.DTYPE_BF16 if (dtype or os.environ.get("SPEAKER_B200_DTYPE", "fp32")) == "bf16" else _native.DTYPE_F32
        self.pool = _native.POOL_MAX if (pool or os.environ.get("SPEAKER_B200_POOL", "mean")) == "max" else _native.POOL_MEAN
        self.k = max(1, min(_native.MAX_K, int(k or os.environ.get("SPEAKER_B200_TOPK", 10))))
        self.threshold = float(threshold)
        self.bank: Optional[store.Bank] = None

    def close(self):
        self.ctx.close()

    def load_bank(self, tags: Optional[str] = None, use_cache: bool = True) -> store.Bank:
        speakers = store.list_all_speakers_cached()
        if tags:
            speakers = store.filter_speakers_by_tags(speakers, [t.strip() for t in tags.split(",")], any_tag=False)
        candidates = [s for s in speakers if s.get("embeddings", {}).get(self.backend_name)]
        build = store.build_bank_cached if use_cache else store.build_bank
        self.bank = build(candidates, self.backend_name)
        if self.bank.P:
            if self.world > 1:
                # this rank's slice of the rows; ids in the results are GLOBAL rows, so the tables above stay whole.
                # An empty slice (more ranks than speakers) is legal: the rank still joins every all-gather.
                from . import sharding
                p0, p1 = sharding.shard_bank_rows(self.bank.row_speaker, self.world)[self.rank]
                self.ctx.bank_load(self.bank.rows[p0:p1].reshape(p1 - p0, self.bank.rows.shape[1]), self.bank.row_speaker[p0:p1],
                                   self.bank.row_trust[p0:p1], dtype=self.dtype, global_row_offset=p0)
            else:
                self.ctx.bank_load(self.bank.rows, self.bank.row_speaker, self.bank.row_trust, dtype=self.dtype)
        return self.bank

    def identify(self, audio_paths: Sequence, assign_threshold: float = 0.3, min_trust: str = "low",
                 expected: Optional[Sequence] = None) -> List[RecordingResult]:
        """expected[i] = (context_name, [expected speaker ids]) of recording i (speaker-assign:531-541) or None."""
        if self.bank is None:
            self.load_bank()
        bank = self.bank
        recs = [store.load_segment_embeddings(p, self.backend_name) for p in audio_paths]
        if not bank.P or not recs:
            return [RecordingResult(Path(p), r.labels, {l: [] for l in r.labels}, {}) for p, r in zip(audio_paths, recs)]
        D = bank.rows.shape[1]
        for p, r in zip(audio_paths, recs):
            if r.emb.shape[1] != D:
                raise ValueError(f"{p}: segment embeddings are {r.emb.shape[1]}-d but the enrolled bank is {D}-d")
        # label groups: recordings back to back (ragged: each recording contributes exactly its own labels)
        goffs = np.cumsum([0] + [len(r.labels) for r in recs])
        # fp16 sidecars stay fp16 all the way to the device (half the PCIe bytes; K1 widens them exactly)
        seg_dtype = np.float16 if all(r.emb.dtype == np.float16 for r in recs) else np.float32
        seg = np.concatenate([np.asarray(r.emb, seg_dtype) for r in recs], axis=0)
        lab = np.concatenate([r.label_index + goffs[i] for i, r in enumerate(recs)]).astype(np.int32)
        L = int(goffs[-1])
        rows, scores, counts = self.ctx.identify(seg, lab, L, pool=self.pool, threshold=self.threshold, k=self.k)
        self.ctx.assign(assign_threshold, min_trust)
        out = self.ctx.fetch(with_assign=True)
        trust_names = _native.TRUST_NAMES
        results = []
        for i, (p, r) in enumerate(zip(audio_paths, recs)):
            matches, mappings = {}, {}
            context_name, expected_speakers = (expected[i] if expected and expected[i] else (None, None))
            for j, label in enumerate(r.labels):
                g = int(goffs[i]) + j
                lst = []
                for rank in range(int(counts[g])):
                    br = int(rows[g, rank])
                    sim = float(scores[g, rank])
                    lst.append({"speaker_id": bank.speaker_ids[int(bank.row_speaker[br])], "score": sim, "confidence": sim,
                                "trust_level": trust_names[int(out["trust"][g, rank])], "embedding_id": bank.row_emb_id[br],
                                "backend": self.backend_name, "label": label, "rank": rank})
                matches[label] = lst
                if expected_speakers:
                    # context signals take part: general fusion in Python (speaker-assign:418-492 restated in signals.py)
                    sigs = sg.signals_from_matches(lst, min_trust=min_trust, label=label)
                    sigs += sg.collect_context_signals(label, context_name, expected_speakers)
                    a = sg.combine_signals(label, sigs, threshold=assign_threshold)
                    m = {"speaker_id": a.speaker_id, "confidence": a.confidence, "score": round(a.score, 3), "signals": a.signals}
                    if a.candidates:
                        m["candidates"] = a.candidates
                else:
                    # embeddings are the only signal: the device already ran combine_signals (sdk_assign)
                    ai = int(out["assign_idx"][g])
                    kept = [row for row in lst if _passes(row["trust_level"], min_trust)]
                    chosen = lst[ai] if ai >= 0 else None
                    best = _best_index(out, g)
                    ev_row = lst[best] if best is not None else None
                    m = {"speaker_id": chosen["speaker_id"] if chosen else None,
                         "confidence": _native.CONF_NAMES[int(out["assign_conf"][g])],
                         "score": round(float(out["assign_score"][g]), 3),
                         "signals": ([{"type": "embedding_match", "score": ev_row["score"], "embedding_id": ev_row["embedding_id"],
                                       "trust_level": ev_row["trust_level"], "backend": ev_row["backend"]}] if (kept and ev_row) else [])}
                    cands = [{"speaker_id": lst[int(ci)]["speaker_id"], "score": float(cs)}
                             for ci, cs in zip(out["cand_idx"][g], out["cand_score"][g]) if ci >= 0]
                    if cands:
                        m["candidates"] = cands
                mappings[label] = m
            results.append(RecordingResult(Path(p), r.labels, matches, mappings))
        return results


def _passes(trust: str, min_trust: str) -> bool:
    order = ["low", "medium", "high"]
    if min_trust in order and trust in order:
        return order.index(trust) >= order.index(min_trust)
    return True


def _best_index(out, g: int) -> Optional[int]:
    """Index of the match whose evidence combine_signals reports: the assigned one, else the first candidate
    (unassigned case: candidates start with the best, speaker-assign:475-483)."""
    ai = int(out["assign_idx"][g])
    if ai >= 0:
        return ai
    c0 = int(out["cand_idx"][g, 0])
    return c0 if c0 >= 0 else None
