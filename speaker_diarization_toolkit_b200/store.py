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
import gc
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


def list_all_speakers_cached() -> List[Dict[str, Any]]:
    """Same list as `list_all_speakers`, without re-parsing a million unchanged JSON files on every call (SURVEY 8f item 1:
    the reference's per-call O(P) pass, speaker_detection:206-220, is what a million-profile identify waits for).
    `db/.profiles.pack` holds the parsed profiles in file order plus every file's (size, mtime_ns); one scandir + stat pass
    finds what changed, only those files are parsed again, and the pack is rewritten when anything did.  The per-file
    layout stays authoritative.  SPEAKER_B200_PROFILE_CACHE=0 disables it."""
    db = get_speakers_db_path()
    if os.environ.get("SPEAKER_B200_PROFILE_CACHE", "1") == "0" or not db.exists():
        return list_all_speakers()
    pack_path = db / ".profiles.pack"
    cur = {}
    with os.scandir(db) as it:
        for e in it:
            n = e.name
            if n.endswith(".json") and n[0] != ".":
                st = e.stat()
                cur[n] = [st.st_size, st.st_mtime_ns]
    names = sorted(cur)
    old_stat, old_prof = {}, {}
    if pack_path.exists():
        try:
            gc_was_on = gc.isenabled()
            gc.disable()                       # (a million fresh dicts: the cycle collector would walk them again and again)
            try:
                with open(pack_path, "r") as fh:
                    pack = json.load(fh)
            finally:
                if gc_was_on:
                    gc.enable()
            # nothing changed (the usual call): two list comparisons, no per-profile Python work
            if pack["names"] == names and pack["stats"] == [cur[n] for n in names] and len(pack["profiles"]) == len(names):
                return pack["profiles"]
            old_stat = dict(zip(pack["names"], pack["stats"]))
            old_prof = dict(zip(pack["names"], pack["profiles"]))
        except (OSError, ValueError, KeyError, TypeError):
            old_stat, old_prof = {}, {}
    out: List[Dict[str, Any]] = []
    kept_names, kept_stats = [], []
    dirty = len(old_stat) != len(cur)
    for n in names:
        if old_stat.get(n) == cur[n]:
            prof = old_prof[n]
        else:
            dirty = True
            try:
                with open(db / n, "r") as fh:
                    prof = json.load(fh)
            except (json.JSONDecodeError, IOError) as exc:
                print(f"Warning: Failed to load {db / n}: {exc}", file=sys.stderr)
                continue
        out.append(prof)
        kept_names.append(n)
        kept_stats.append(cur[n])
    if dirty:
        try:
            tmp = pack_path.with_name(pack_path.name + f".{os.getpid()}.tmp")
            with open(tmp, "w") as fh:
                json.dump({"names": kept_names, "stats": kept_stats, "profiles": out}, fh)
            os.replace(tmp, pack_path)
        except OSError as exc:
            print(f"Warning: could not update the profile pack: {exc}", file=sys.stderr)
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
    emb: np.ndarray          # [N, D] fp32 (or fp16: compact sidecar), sorted by label (stable)
    label_index: np.ndarray  # [N] int32 index into labels, non-decreasing
    labels: List[str]        # sorted(set(labels))  (speaker-assign:196)
    start: np.ndarray
    end: np.ndarray


def save_segment_embeddings(audio_path, backend_name: str, emb, labels, start=None, end=None, dtype=None) -> Path:
    """`dtype` = np.float16 stores the compact form of the sidecar (half the bytes on disk and over PCIe; the device
    widens every element exactly to fp32 before the canonical normalise); default: float32, or float16 if `emb`
    already is.  Labels keep their full length (the reference puts no limit on diarization labels)."""
    emb = np.asarray(emb)
    want = np.dtype(dtype) if dtype is not None else (np.dtype(np.float16) if emb.dtype == np.float16 else np.dtype(np.float32))
    if want not in (np.dtype(np.float16), np.dtype(np.float32)):
        raise ValueError("segment embeddings are stored as float32 or float16")
    emb = np.ascontiguousarray(emb, dtype=want)
    n = emb.shape[0]
    path = sidecar_path(audio_path, backend_name)
    np.savez(path, emb=emb, label=np.asarray([str(x) for x in labels], dtype=str),
             start=np.zeros(n) if start is None else np.asarray(start, np.float64),
             end=np.zeros(n) if end is None else np.asarray(end, np.float64))
    return path


def sidecar_segment_count(audio_path, backend_name: str) -> int:
    """Number of segments in a sidecar without reading the embeddings (partitioning a manifest over GPUs)."""
    with np.load(sidecar_path(audio_path, backend_name), allow_pickle=False) as z:
        return int(z["label"].shape[0])


def load_segment_embeddings(audio_path, backend_name: str) -> SegmentEmbeddings:
    path = sidecar_path(audio_path, backend_name)
    if not path.exists():
        raise FileNotFoundError(f"segment-embedding sidecar not found: {path}")
    with np.load(path, allow_pickle=False) as z:
        raw = z["emb"]
        emb = np.ascontiguousarray(raw, dtype=np.float16 if raw.dtype == np.float16 else np.float32)
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
# `embeddings/.bank-<backend>-D<d>.idx.npz` (BINARY index: key "<speaker>/<emb id>" -> row, source file size +
# mtime_ns; numpy arrays, so a million entries load in a fraction of a second where JSON took several).  A row is
# reused only if its source file's (size, mtime_ns) still match; new / changed files are re-read and appended; the
# pack is rewritten compacted when more than a quarter of it is stale.  Everything per-row is vectorised: one dict
# lookup pass for key -> row, NumPy comparisons for the stat check, ONE fancy-index gather from the mmap (or the mmap
# itself when the pack already is exactly the wanted rows in order).
# SPEAKER_B200_BANK_CACHE=trust skips the per-file stat() calls (a pack row is reused whenever its key is present):
# for stores whose vectors are never rewritten in place.
def _cache_paths(backend_name: str, dim: int):
    root = get_embeddings_path()
    return root / f".bank-{backend_name}-D{dim}.f32", root / f".bank-{backend_name}-D{dim}.idx.npz"


def _load_cache_index(pack_path: Path, idx_path: Path, dim: int):
    """(key -> position dict, row [n], size [n], mtime_ns [n], memmap) or None."""
    if not (pack_path.exists() and idx_path.exists()):
        return None
    try:
        with np.load(idx_path, allow_pickle=False) as z:
            if int(z["dim"]) != dim:
                return None
            keys, row, size, mtime, n_rows = z["key"], z["row"], z["size"], z["mtime_ns"], int(z["rows"])
        if pack_path.stat().st_size != n_rows * dim * 4 or not n_rows:
            return None
        pack = np.memmap(pack_path, dtype=np.float32, mode="r", shape=(n_rows, dim))
        return dict(zip(keys.tolist(), range(len(keys)))), row, size, mtime, pack
    except (OSError, ValueError, KeyError):
        return None


def _write_cache_index(idx_path: Path, dim: int, backend_name: str, keys, row, size, mtime, n_rows: int) -> None:
    tmp = idx_path.with_name(idx_path.name + f".{os.getpid()}.tmp.npz")
    np.savez(tmp, dim=np.int64(dim), rows=np.int64(n_rows), backend=np.asarray(backend_name), key=np.asarray(keys, dtype=str),
             row=np.asarray(row, np.int64), size=np.asarray(size, np.int64), mtime_ns=np.asarray(mtime, np.int64))
    os.replace(tmp, idx_path)


def build_bank_cached(candidates: List[Dict[str, Any]], backend_name: str, dim: Optional[int] = None) -> Bank:
    trust_mode = os.environ.get("SPEAKER_B200_BANK_CACHE", "") == "trust"
    root = get_embeddings_path()
    unknown = TRUST_CODES["unknown"]
    keys: List[str] = []
    paths: List[Any] = []           # str / Path of the vector file, None = not resolved yet (trust mode)
    sizes: List[int] = []
    mtimes: List[int] = []
    row_speaker_l: List[int] = []
    row_trust_l: List[int] = []
    emb_ids: List[Optional[str]] = []
    spk_recs = []       # (speaker id, record) of every row, for lazy path resolution
    speaker_ids: List[str] = []
    root_s = str(root)
    for prof in candidates:
        sid = prof.get("id")
        recs = (prof.get("embeddings") or {}).get(backend_name) or []
        if not recs:
            continue
        idx = len(speaker_ids)
        used = False
        for rec in recs:
            rid = rec.get("id")
            path = None
            if not trust_mode:
                # one stat() per record on the canonical location (existence, size and mtime in a single system call);
                # the other locations (content-addressed handle, explicit file) only if that one is missing
                st = None
                if rid:
                    path = f"{root_s}/{sid}/{rid}.npy"
                    try:
                        st = os.stat(path)
                    except OSError:
                        st = None
                if st is None:
                    path = vector_path(sid, rec)
                    if path is None:
                        print(f"Warning: no vector file for {sid}/{rid} ({backend_name})", file=sys.stderr)
                        continue
                    st = path.stat()
                sizes.append(st.st_size)
                mtimes.append(st.st_mtime_ns)
            used = True
            keys.append(f"{sid}/{rid}")
            paths.append(path)
            spk_recs.append((sid, rec))
            row_speaker_l.append(idx)
            row_trust_l.append(TRUST_CODES.get(rec.get("trust_level", "unknown"), unknown))
            emb_ids.append(rid)
        if used:                        # (a speaker without a usable record does not consume an index)
            speaker_ids.append(sid)
    n = len(keys)
    if not n:
        return Bank(np.zeros((0, dim or 0), np.float32), np.zeros(0, np.int32), np.zeros(0, np.uint8), [], [])
    cache = None
    if dim is None:
        # The dimension of the WANTED vectors (not of whatever pack happens to be newest: a store re-enrolled at another D,
        # or tag-filtered sub-banks of different D, must not inherit a stale pack's dimension): a pack that holds the
        # first wanted key with an unchanged source file tells it without opening a .npy; otherwise that file's header.
        st0 = None if trust_mode else os.stat(paths[0])
        for ip in sorted(root.glob(f".bank-{backend_name}-D*.idx.npz")):
            try:
                d = int(ip.name.split("-D")[-1].split(".")[0])
            except ValueError:
                continue
            cand = _load_cache_index(_cache_paths(backend_name, d)[0], ip, d)
            if cand is None:
                continue
            at = cand[0].get(keys[0])
            if at is not None and (trust_mode or (cand[2][at] == st0.st_size and cand[3][at] == st0.st_mtime_ns)):
                dim, cache = d, cand
                break
        if dim is None:
            first = paths[0] if paths[0] is not None else vector_path(*spk_recs[0])
            if first is None:
                raise ValueError(f"no vector file for {keys[0]} ({backend_name})")
            dim = int(np.load(first, mmap_mode="r").reshape(-1).shape[0])
    pack_path, idx_path = _cache_paths(backend_name, dim)
    if cache is None:
        cache = _load_cache_index(pack_path, idx_path, dim)
    size, mtime = np.zeros(n, np.int64), np.zeros(n, np.int64)
    if not trust_mode:
        size, mtime = np.asarray(sizes, np.int64), np.asarray(mtimes, np.int64)
    hit = np.zeros(n, bool)
    src = np.zeros(n, np.int64)
    kpos, pack = {}, None
    if cache is not None:
        kpos, c_row, c_size, c_mtime, pack = cache
        pos = np.fromiter((kpos.get(k, -1) for k in keys), np.int64, n)
        hit = pos >= 0
        pc = np.where(hit, pos, 0)
        if trust_mode:                                   # rows taken on trust keep the stat the index already holds
            size, mtime = np.where(hit, c_size[pc], 0), np.where(hit, c_mtime[pc], 0)
        else:
            hit &= (c_size[pc] == size) & (c_mtime[pc] == mtime)
        src = np.where(hit, c_row[pc], 0)
        hit &= src < pack.shape[0]
    if pack is not None and hit.all() and n == pack.shape[0] and np.array_equal(src, np.arange(n)):
        rows = np.asarray(pack)                          # the pack IS the bank: no copy, pages come in as they are read
    else:
        rows = np.empty((n, dim), dtype=np.float32)
        if pack is not None and hit.any():
            rows[hit] = pack[src[hit]]                   # one gather
    miss = np.flatnonzero(~hit)
    for i in miss.tolist():
        path = paths[i] if paths[i] is not None else vector_path(*spk_recs[i])
        if path is None:
            raise ValueError(f"no vector file for {keys[i]} ({backend_name})")
        vec = np.load(path).astype(np.float32).reshape(-1)
        if vec.shape[0] != dim:
            raise ValueError(f"embedding {keys[i]} has dimension {vec.shape[0]}, expected {dim}")
        rows[i] = vec
        if trust_mode:
            stt = os.stat(path)
            size[i], mtime[i] = stt.st_size, stt.st_mtime_ns
    if len(miss) or pack is None:
        try:
            live = set(keys)
            stale = sum(1 for k in kpos if k not in live)
            if pack is None or stale * 4 > max(1, len(kpos)):
                # rewrite compacted: exactly the rows in use now (unique temporary name: two processes rewriting at once
                # must not interleave their rows)
                tmp = pack_path.with_name(pack_path.name + f".{os.getpid()}.tmp")
                np.ascontiguousarray(rows).tofile(tmp)
                os.replace(tmp, pack_path)
                _write_cache_index(idx_path, dim, backend_name, keys, np.arange(n), size, mtime, n)
            else:
                # append the new / changed rows
                n_rows = pack.shape[0]
                del pack
                with open(pack_path, "ab") as fh:
                    np.ascontiguousarray(rows[miss]).tofile(fh)
                k_row, k_size, k_mtime = (np.asarray(x, np.int64).copy() for x in (c_row, c_size, c_mtime))
                add_k, add_r, add_s, add_m = [], [], [], []
                for j, i in enumerate(miss.tolist()):
                    at = kpos.get(keys[i])
                    if at is not None:
                        k_row[at], k_size[at], k_mtime[at] = n_rows + j, size[i], mtime[i]
                    else:
                        add_k.append(keys[i]); add_r.append(n_rows + j); add_s.append(size[i]); add_m.append(mtime[i])
                _write_cache_index(idx_path, dim, backend_name, list(kpos.keys()) + add_k, np.r_[k_row, np.asarray(add_r, np.int64)],
                                   np.r_[k_size, np.asarray(add_s, np.int64)], np.r_[k_mtime, np.asarray(add_m, np.int64)],
                                   n_rows + len(miss))
        except OSError as exc:      # read-only store: the cache is an optimisation, never a requirement
            print(f"Warning: could not update the packed bank cache: {exc}", file=sys.stderr)
    return Bank(rows, np.asarray(row_speaker_l, np.int32), np.asarray(row_trust_l, np.uint8), emb_ids, speaker_ids)
