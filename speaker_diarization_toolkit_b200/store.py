"""Profile / embedding storage under $SPEAKERS_EMBEDDINGS_DIR -- read side of the hot path.

Layout (unchanged from the reference):
    db/<speaker_id>.json                     profile, `embeddings[backend] = [record, ...]`  (speaker_detection:155-220)
    embeddings/<speaker_id>/<emb_id>.npy     one fp32 [D] vector per embedding record        (base.py:123,
                                             ramblings/2026-01-13--speaker-identity-system.md:115-118)
    <audio>.<backend>.segemb.npz             per-segment embeddings of a recording (sidecar; the embedding
                                             extractors of the reference are network APIs)
`build_bank` turns the candidates list that `identify_speaker` receives (base.py:133) into the contiguous
[P, D] fp32 bank + row->speaker / row->trust / row->embedding-id tables that the C-ABI loads.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from ._native import TRUST_CODES

DEFAULT_DB_DIR = os.path.expanduser("~/.config/speakers_embeddings")


def get_db_dir() -> Path:
    return Path(os.environ.get("SPEAKERS_EMBEDDINGS_DIR", DEFAULT_DB_DIR))


def get_speakers_db_path() -> Path:
    return get_db_dir() / "db"


def get_embeddings_path() -> Path:
    return get_db_dir() / "embeddings"


def load_speaker(speaker_id: str) -> Optional[Dict[str, Any]]:
    path = get_speakers_db_path() / f"{speaker_id}.json"
    if not path.exists():
        return None
    with open(path, "r") as fh:
        return json.load(fh)


def normalize_speaker_id(speaker_id: str) -> str:
    """speaker_detection:146-148."""
    return speaker_id.lower().replace(" ", "-")


def save_speaker(profile: Dict[str, Any]) -> None:
    """speaker_detection:188-194 (stamps updated_at, 2-space JSON)."""
    from datetime import datetime, timezone
    get_speakers_db_path().mkdir(parents=True, exist_ok=True)
    profile["updated_at"] = datetime.now(timezone.utc).isoformat()
    with open(get_speakers_db_path() / f"{profile['id']}.json", "w") as fh:
        json.dump(profile, fh, indent=2, ensure_ascii=False)


def get_samples_by_source_audio(speaker_id: str, audio_b3sum: str) -> Dict[str, List[str]]:
    """Sample hashes of one source recording by review status (speaker_detection:310-356): reads
    samples/<speaker_id>/*.meta.yaml written by `speaker_samples`; no samples directory -> three empty lists."""
    result: Dict[str, List[str]] = {"reviewed": [], "unreviewed": [], "rejected": []}
    sdir = get_db_dir() / "samples" / speaker_id
    if not sdir.exists():
        return result
    try:
        import yaml
    except ImportError:
        yaml = None
    for meta_path in sorted(sdir.glob("*.meta.yaml")):
        try:
            text = meta_path.read_text()
            meta = (yaml.safe_load(text) if yaml else json.loads(text)) or {}
        except Exception:
            continue
        if meta.get("source", {}).get("audio_b3sum") != audio_b3sum or not meta.get("b3sum"):
            continue
        status = meta.get("review", {}).get("status", "pending")
        key = "reviewed" if status == "reviewed" else "rejected" if status == "rejected" else "unreviewed"
        result[key].append(meta["b3sum"])
    return result


def list_all_speakers() -> List[Dict[str, Any]]:
    """Every db/*.json in sorted file order (speaker_detection:206-220); unreadable files are skipped with a warning."""
    db = get_speakers_db_path()
    out: List[Dict[str, Any]] = []
    if not db.exists():
        return out
    for path in sorted(db.glob("*.json")):
        try:
            with open(path, "r") as fh:
                out.append(json.load(fh))
        except (json.JSONDecodeError, IOError) as exc:
            print(f"Warning: Failed to load {path}: {exc}", file=sys.stderr)
    return out


def filter_speakers_by_tags(speakers, tags: Optional[Sequence[str]] = None, any_tag: bool = False):
    """AND (default) / OR tag filter (speaker_detection:223-246)."""
    if not tags:
        return speakers
    want = set(tags)
    keep = []
    for spk in speakers:
        have = set(spk.get("tags", []))
        if (have & want) if any_tag else (want <= have):
            keep.append(spk)
    return keep


def compute_trust_level(samples: Dict[str, List[str]]) -> str:
    """speaker_detection:359-379."""
    if samples.get("rejected", []):
        return "invalidated"
    reviewed, unreviewed = samples.get("reviewed", []), samples.get("unreviewed", [])
    if reviewed and not unreviewed:
        return "high"
    if reviewed:
        return "medium"
    return "low"


# ---- vectors ------------------------------------------------------------------------------------
def vector_path(speaker_id: str, record: Dict[str, Any]) -> Optional[Path]:
    """Where the fp32 vector of an embedding record lives.  Order: the canonical per-speaker file, the
    content-addressed handle that `enroll_speaker` returned as `external_id` (the reference CLI persists
    only that key, speaker_detection:890-911), an explicit `file` key."""
    root = get_embeddings_path()
    cands = []
    if record.get("id"):
        cands.append(root / speaker_id / f"{record['id']}.npy")
    ext = record.get("external_id")
    if isinstance(ext, str) and ext.endswith(".npy"):
        cands.append(root / ext)
    if record.get("file"):
        f = Path(record["file"])
        cands.append(f if f.is_absolute() else get_db_dir() / f)
    for c in cands:
        if c.exists():
            return c
    return None


def bank_dimension(backend_name: str, speakers: Optional[List[Dict[str, Any]]] = None) -> Optional[int]:
    """Dimension of the vectors already enrolled for `backend_name` (first readable one), or None for an empty bank."""
    for prof in (speakers if speakers is not None else list_all_speakers()):
        for rec in (prof.get("embeddings") or {}).get(backend_name) or []:
            path = vector_path(prof.get("id"), rec)
            if path is not None:
                try:
                    return int(np.load(path, mmap_mode="r").reshape(-1).shape[0])
                except (OSError, ValueError):
                    continue
    return None


def store_vector_canonical(speaker_id: str, emb_id: str, vec: np.ndarray) -> Path:
    """embeddings/<speaker_id>/<emb_id>.npy -- the authoritative per-file layout (module docstring)."""
    dst = get_embeddings_path() / speaker_id / f"{emb_id}.npy"
    dst.parent.mkdir(parents=True, exist_ok=True)
    np.save(dst, np.ascontiguousarray(vec, dtype=np.float32))
    return dst


def store_vector_cas(vec: np.ndarray) -> str:
    """Content-addressed write; returns the handle relative to embeddings/ (used as `external_id`)."""
    vec = np.ascontiguousarray(vec, dtype=np.float32)
    digest = hashlib.sha256(vec.tobytes()).hexdigest()[:32]
    rel = Path("_cas") / digest[:2] / f"{digest}.npy"
    dst = get_embeddings_path() / rel
    dst.parent.mkdir(parents=True, exist_ok=True)
    if not dst.exists():
        np.save(dst, vec)
    return str(rel)


@dataclass
class Bank:
    rows: np.ndarray            # [P, D] fp32 raw vectors
    row_speaker: np.ndarray     # [P] int32 index into speaker_ids (contiguous per speaker)
    row_trust: np.ndarray       # [P] uint8
    row_emb_id: List[Optional[str]]
    speaker_ids: List[str]

    @property
    def P(self) -> int:
        return self.rows.shape[0]


def build_bank(candidates: List[Dict[str, Any]], backend_name: str, dim: Optional[int] = None) -> Bank:
    """candidates: speaker profiles (dicts, NOT mutated) each with embeddings[backend_name] records.
    Rows follow candidate order then record order, so a speaker's rows are contiguous."""
    rows, row_speaker, row_trust, row_emb, speaker_ids = [], [], [], [], []
    for prof in candidates:
        sid = prof.get("id")
        records = (prof.get("embeddings") or {}).get(backend_name) or []
        idx = None
        for rec in records:
            path = vector_path(sid, rec)
            if path is None:
                print(f"Warning: no vector file for {sid}/{rec.get('id')} ({backend_name})", file=sys.stderr)
                continue
            vec = np.load(path).astype(np.float32).reshape(-1)
            if dim is None:
                dim = vec.shape[0]
            if vec.shape[0] != dim:
                raise ValueError(f"embedding {sid}/{rec.get('id')} has dimension {vec.shape[0]}, expected {dim}")
            if idx is None:
                idx = len(speaker_ids)
                speaker_ids.append(sid)
            rows.append(vec)
            row_speaker.append(idx)
            row_trust.append(TRUST_CODES.get(rec.get("trust_level", "unknown"), TRUST_CODES["unknown"]))
            row_emb.append(rec.get("id"))
    if not rows:
        return Bank(np.zeros((0, dim or 0), np.float32), np.zeros(0, np.int32), np.zeros(0, np.uint8), [], [])
    return Bank(np.stack(rows), np.asarray(row_speaker, np.int32), np.asarray(row_trust, np.uint8), row_emb, speaker_ids)


# ---- per-segment embedding sidecar ---------------------------------------------------------------
def sidecar_path(audio_path, backend_name: str) -> Path:
    return Path(f"{audio_path}.{backend_name}.segemb.npz")


@dataclass
class SegmentEmbeddings:
    emb: np.ndarray          # [N, D] fp32, sorted by label (stable)
    label_index: np.ndarray  # [N] int32 index into labels, non-decreasing
    labels: List[str]        # sorted(set(labels))  (speaker-assign:196)
    start: np.ndarray
    end: np.ndarray


def save_segment_embeddings(audio_path, backend_name: str, emb, labels, start=None, end=None) -> Path:
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    n = emb.shape[0]
    path = sidecar_path(audio_path, backend_name)
    np.savez(path, emb=emb, label=np.asarray(labels, dtype="<U32"),
             start=np.zeros(n) if start is None else np.asarray(start, np.float64),
             end=np.zeros(n) if end is None else np.asarray(end, np.float64))
    return path


def load_segment_embeddings(audio_path, backend_name: str) -> SegmentEmbeddings:
    path = sidecar_path(audio_path, backend_name)
    if not path.exists():
        raise FileNotFoundError(f"segment-embedding sidecar not found: {path}")
    with np.load(path, allow_pickle=False) as z:
        emb = np.ascontiguousarray(z["emb"], dtype=np.float32)
        lab = [str(x) for x in z["label"]]
        start = np.asarray(z["start"], np.float64) if "start" in z else np.zeros(len(lab))
        end = np.asarray(z["end"], np.float64) if "end" in z else np.zeros(len(lab))
    if emb.ndim != 2 or emb.shape[0] != len(lab):
        raise ValueError(f"bad sidecar {path}: emb {emb.shape} vs {len(lab)} labels")
    labels = sorted(set(lab))
    pos = {l: i for i, l in enumerate(labels)}
    idx = np.asarray([pos[l] for l in lab], dtype=np.int32)
    order = np.argsort(idx, kind="stable")
    return SegmentEmbeddings(emb[order], idx[order], labels, start[order], end[order])


# ---- packed bank cache ------------------------------------------------------------
# `cmd_identify` would otherwise open one .npy per embedding record on every call (speaker_detection:1054 ->
# base.py:123); at a million-profile bank that, not the GPU, is the wall time.  The per-file layout stays
# authoritative: the cache is `embeddings/.bank-<backend>-D<d>.f32` (raw fp32 [rows, D], mmap-able) +
# `embeddings/.bank-<backend>-D<d>.idx.json` (speaker/emb id -> row, file size + mtime_ns).  A row is reused only if
# its source file's (size, mtime_ns) still match; new / changed files are re-read and appended; the pack is
# rewritten compacted when more than a quarter of it is stale.
def _cache_paths(backend_name: str, dim: int):
    root = get_embeddings_path()
    return root / f".bank-{backend_name}-D{dim}.f32", root / f".bank-{backend_name}-D{dim}.idx.json"


def build_bank_cached(candidates: List[Dict[str, Any]], backend_name: str, dim: Optional[int] = None) -> Bank:
    wanted = []     # (speaker index, speaker id, record, path, stat key)
    speaker_ids: List[str] = []
    for prof in candidates:
        sid = prof.get("id")
        idx = None
        for rec in (prof.get("embeddings") or {}).get(backend_name) or []:
            path = vector_path(sid, rec)
            if path is None:
                print(f"Warning: no vector file for {sid}/{rec.get('id')} ({backend_name})", file=sys.stderr)
                continue
            if idx is None:
                idx = len(speaker_ids)
                speaker_ids.append(sid)
            st = path.stat()
            wanted.append((idx, sid, rec, path, [str(path), st.st_size, st.st_mtime_ns]))
    if not wanted:
        return Bank(np.zeros((0, dim or 0), np.float32), np.zeros(0, np.int32), np.zeros(0, np.uint8), [], [])
    if dim is None:
        packs = sorted(get_embeddings_path().glob(f".bank-{backend_name}-D*.idx.json"), key=lambda q: q.stat().st_mtime_ns)
        if packs:       # a pack exists: its dimension (a mismatching vector is caught when it is read)
            dim = int(packs[-1].name.split("-D")[-1].split(".")[0])
        else:
            dim = int(np.load(wanted[0][3]).reshape(-1).shape[0])
    pack_path, idx_path = _cache_paths(backend_name, dim)
    index, pack = {}, None
    if pack_path.exists() and idx_path.exists():
        try:
            meta = json.loads(idx_path.read_text())
            if meta.get("dim") == dim and pack_path.stat().st_size == meta.get("rows", -1) * dim * 4:
                index = {k: v for k, v in meta.get("entries", {}).items()}
                pack = np.memmap(pack_path, dtype=np.float32, mode="r", shape=(meta["rows"], dim)) if meta["rows"] else None
        except (json.JSONDecodeError, OSError, ValueError):
            index, pack = {}, None
    rows = np.empty((len(wanted), dim), dtype=np.float32)
    fresh = []      # (position, key, stat) that had to be read from the per-file layout
    for pos, (_, sid, rec, path, stat) in enumerate(wanted):
        key = f"{sid}/{rec.get('id')}"
        ent = index.get(key)
        if ent is not None and pack is not None and ent["stat"] == stat and ent["row"] < pack.shape[0]:
            rows[pos] = pack[ent["row"]]
            continue
        vec = np.load(path).astype(np.float32).reshape(-1)
        if vec.shape[0] != dim:
            raise ValueError(f"embedding {key} has dimension {vec.shape[0]}, expected {dim}")
        rows[pos] = vec
        fresh.append((pos, key, stat))
    if fresh or pack is None:
        live = {f"{sid}/{rec.get('id')}" for _, sid, rec, _, _ in wanted}
        stale = sum(1 for k in index if k not in live)
        try:
            if pack is None or stale * 4 > max(1, len(index)):
                # rewrite compacted: exactly the rows in use now
                tmp = pack_path.with_suffix(".tmp")
                rows.tofile(tmp)
                os.replace(tmp, pack_path)
                entries = {f"{sid}/{rec.get('id')}": {"row": pos, "stat": stat} for pos, (_, sid, rec, _, stat) in enumerate(wanted)}
                n_rows = len(wanted)
            else:
                # append the new / changed rows
                n_rows = pack.shape[0]
                del pack
                with open(pack_path, "ab") as fh:
                    for pos, key, stat in fresh:
                        rows[pos].tofile(fh)
                        index[key] = {"row": n_rows, "stat": stat}
                        n_rows += 1
                entries = index
            tmpi = idx_path.with_suffix(".tmp")
            tmpi.write_text(json.dumps({"dim": dim, "rows": n_rows, "backend": backend_name, "entries": entries}))
            os.replace(tmpi, idx_path)
        except OSError as exc:      # read-only store: the cache is an optimisation, never a requirement
            print(f"Warning: could not update the packed bank cache: {exc}", file=sys.stderr)
    row_speaker = np.asarray([w[0] for w in wanted], np.int32)
    row_trust = np.asarray([TRUST_CODES.get(w[2].get("trust_level", "unknown"), TRUST_CODES["unknown"]) for w in wanted], np.uint8)
    return Bank(rows, row_speaker, row_trust, [w[2].get("id") for w in wanted], speaker_ids)
