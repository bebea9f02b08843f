"""ctypes wrapper around oracle/canonical.c (TEST INFRASTRUCTURE -- see the header of that file).

Builds `oracle/_build/liboracle.so` with gcc on first use.  Importers: tests/, __graft_entry__.smoke(),
bench.py (cpu_baseline leg only).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "_build" / "liboracle.so"
    src = _HERE / "canonical.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s", "all"], check=True,
                       capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_bf16_round.restype = C.c_float
        _LIB.orc_bf16_round.argtypes = [C.c_float]
        _LIB.orc_pair_q30.restype = C.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def normalize(x, mode: int = 0):
    """Canonical normalise (+bf16 rounding when mode == 1).  Returns (operands fp32 [n,D], norms [n])."""
    x = _f32(x)
    n, D = x.shape
    out = np.empty_like(x)
    norms = np.empty(n, dtype=np.float32)
    lib().orc_normalize(_p(x, C.c_float), C.c_int64(n), C.c_int32(D), C.c_int32(mode),
                        _p(out, C.c_float), _p(norms, C.c_float))
    return out, norms


def pooled(seg_ops, goff, bank_ops, pool: int = 0):
    """Canonical pooled similarity [G,P] from already-normalised operands."""
    seg_ops, bank_ops = _f32(seg_ops), _f32(bank_ops)
    goff = np.ascontiguousarray(goff, dtype=np.int64)
    G, P, D = len(goff) - 1, bank_ops.shape[0], bank_ops.shape[1]
    out = np.zeros((G, P), dtype=np.float32)
    lib().orc_pooled(_p(seg_ops, C.c_float), _p(goff, C.c_int64), C.c_int32(G), _p(bank_ops, C.c_float),
                     C.c_int64(P), C.c_int32(D), C.c_int32(pool), _p(out, C.c_float))
    return out


def select(sim, gcount, row_speaker, n_speakers, threshold, k, row_offset=0):
    sim = _f32(sim)
    G, P = sim.shape
    gcount = np.ascontiguousarray(gcount, dtype=np.int64)
    row_speaker = np.ascontiguousarray(row_speaker, dtype=np.int32)
    out_row = np.empty((G, k), dtype=np.int64)
    out_score = np.empty((G, k), dtype=np.float32)
    out_count = np.empty(G, dtype=np.int32)
    lib().orc_select(_p(sim, C.c_float), _p(gcount, C.c_int64), C.c_int32(G), C.c_int64(P),
                     _p(row_speaker, C.c_int32), C.c_int32(n_speakers), C.c_double(threshold), C.c_int32(k),
                     C.c_int64(row_offset), _p(out_row, C.c_int64), _p(out_score, C.c_float),
                     _p(out_count, C.c_int32))
    return out_row, out_score, out_count


def identify(seg_raw, goff, bank_raw, row_speaker, n_speakers, mode=0, pool=0, threshold=0.354, k=10,
             row_offset=0):
    """Whole path on raw (un-normalised) inputs.  Returns (rows [G,k] int64, scores [G,k] f32, count [G])."""
    seg_raw, bank_raw = _f32(seg_raw), _f32(bank_raw)
    goff = np.ascontiguousarray(goff, dtype=np.int64)
    row_speaker = np.ascontiguousarray(row_speaker, dtype=np.int32)
    G, N, P = len(goff) - 1, seg_raw.shape[0], bank_raw.shape[0]
    D = bank_raw.shape[1]
    out_row = np.empty((G, k), dtype=np.int64)
    out_score = np.empty((G, k), dtype=np.float32)
    out_count = np.empty(G, dtype=np.int32)
    lib().orc_identify(_p(seg_raw, C.c_float), _p(goff, C.c_int64), C.c_int32(G), C.c_int64(N),
                       _p(bank_raw, C.c_float), _p(row_speaker, C.c_int32), C.c_int32(n_speakers),
                       C.c_int64(P), C.c_int32(D), C.c_int32(mode), C.c_int32(pool), C.c_double(threshold),
                       C.c_int32(k), C.c_int64(row_offset), _p(out_row, C.c_int64), _p(out_score, C.c_float),
                       _p(out_count, C.c_int32))
    return out_row, out_score, out_count


def assign(match_row, match_score, match_trust, match_count, assign_threshold=0.3, min_trust_code=2):
    """C restatement of combine_signals over embedding-only signals (speaker-assign:418-492)."""
    match_row = np.ascontiguousarray(match_row, dtype=np.int64)
    match_score = _f32(match_score)
    match_trust = np.ascontiguousarray(match_trust, dtype=np.uint8)
    match_count = np.ascontiguousarray(match_count, dtype=np.int32)
    G, k = match_row.shape
    a_idx = np.empty(G, dtype=np.int32)
    a_score = np.empty(G, dtype=np.float64)
    a_conf = np.empty(G, dtype=np.int32)
    c_idx = np.empty((G, 3), dtype=np.int32)
    c_score = np.empty((G, 3), dtype=np.float64)
    lib().orc_assign(_p(match_row, C.c_int64), _p(match_score, C.c_float), _p(match_trust, C.c_uint8),
                     _p(match_count, C.c_int32), C.c_int32(G), C.c_int32(k), C.c_double(assign_threshold),
                     C.c_int32(min_trust_code), _p(a_idx, C.c_int32), _p(a_score, C.c_double),
                     _p(a_conf, C.c_int32), _p(c_idx, C.c_int32), _p(c_score, C.c_double))
    return a_idx, a_score, a_conf, c_idx, c_score


def affinity(seg_raw, goff, mode=1, pool=0):
    seg_raw = _f32(seg_raw)
    goff = np.ascontiguousarray(goff, dtype=np.int64)
    N, D = seg_raw.shape
    G = len(goff) - 1
    out = np.empty((N, G), dtype=np.float32)
    lib().orc_affinity(_p(seg_raw, C.c_float), _p(goff, C.c_int64), C.c_int32(G), C.c_int64(N), C.c_int32(D),
                       C.c_int32(mode), C.c_int32(pool), _p(out, C.c_float))
    return out
