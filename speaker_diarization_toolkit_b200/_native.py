"""ctypes binding of the C-ABI in include/sdk_b200.h (the only way Python reaches the CUDA kernels).

There is no CPU fallback: if `libsdk_b200.so` is missing, or no sm_100 device is present,
`NativeError` is raised -- loudly -- and nothing else is tried.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import numpy as np

PKG = Path(__file__).resolve().parent
import os as _os

# SDK_B200_LIB overrides the library path (A/B runs of kernel variants on one box)
LIB_PATH = Path(_os.environ.get("SDK_B200_LIB") or PKG / "libsdk_b200.so")

DTYPE_F32, DTYPE_BF16 = 0, 1
POOL_MEAN, POOL_MAX = 0, 1
TRUST_CODES = {"high": 0, "medium": 1, "low": 2, "invalidated": 3, "unknown": 4}
TRUST_NAMES = ["high", "medium", "low", "invalidated", "unknown"]
CONF_NAMES = ["unassigned", "low", "medium", "high"]
MAX_K = 32

# every symbol include/sdk_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "sdk_abi_version", "sdk_device_count", "sdk_nccl_unique_id", "sdk_create", "sdk_destroy", "sdk_last_error",
    "sdk_set_option", "sdk_bank_load", "sdk_bank_load_dev", "sdk_identify", "sdk_identify_dev", "sdk_assign",
    "sdk_results_fetch", "sdk_affinity_pooled", "sdk_affinity_pooled_dev", "sdk_sync", "sdk_stream",
    "sdk_timer_start", "sdk_timer_stop", "sdk_profile_get", "sdk_profile_reset", "sdk_launch_count", "sdk_last_path",
    "sdk_last_retry", "sdk_merge_topk", "sdk_stage_a_fetch", "sdk_identify_f16", "sdk_identify_f16_dev", "sdk_probe_bank_read",
]


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sdk_b200 error {code}: {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree extension.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeError(-2, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.sdk_abi_version.restype = C.c_int
    lib.sdk_device_count.restype = C.c_int
    lib.sdk_nccl_unique_id.argtypes = [vp]
    lib.sdk_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]
    lib.sdk_destroy.argtypes = [vp]
    lib.sdk_destroy.restype = None
    lib.sdk_last_error.argtypes = [vp]
    lib.sdk_last_error.restype = C.c_char_p
    lib.sdk_set_option.argtypes = [vp, C.c_char_p, f64]
    lib.sdk_bank_load.argtypes = [vp, vp, vp, vp, i64, i32, i32, i64]
    lib.sdk_bank_load_dev.argtypes = [vp, vp, vp, vp, i64, i32, i32, i64]
    lib.sdk_identify.argtypes = [vp, vp, vp, i64, i32, i32, f64, i32, vp, vp, vp]
    lib.sdk_identify_dev.argtypes = [vp, vp, vp, i64, i32, i32, f64, i32]
    lib.sdk_identify_f16.argtypes = [vp, vp, vp, i64, i32, i32, f64, i32, vp, vp, vp]
    lib.sdk_identify_f16_dev.argtypes = [vp, vp, vp, i64, i32, i32, f64, i32]
    lib.sdk_assign.argtypes = [vp, f64, i32]
    lib.sdk_results_fetch.argtypes = [vp] + [vp] * 9
    lib.sdk_affinity_pooled.argtypes = [vp, vp, vp, i64, i32, i32, i32, i32, vp, vp]
    lib.sdk_affinity_pooled_dev.argtypes = [vp, vp, vp, i64, i32, i32, i32, i32, vp, vp]
    lib.sdk_probe_bank_read.argtypes = [vp]
    lib.sdk_sync.argtypes = [vp]
    lib.sdk_stream.argtypes = [vp]
    lib.sdk_stream.restype = vp
    lib.sdk_timer_start.argtypes = [vp]
    lib.sdk_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    lib.sdk_profile_get.argtypes = [vp, C.c_char_p, C.POINTER(C.c_float), C.POINTER(i64)]
    lib.sdk_profile_reset.argtypes = [vp]
    lib.sdk_launch_count.argtypes = [vp]
    lib.sdk_launch_count.restype = i64
    lib.sdk_last_path.argtypes = [vp, C.POINTER(i32), C.POINTER(i64)]
    lib.sdk_last_retry.argtypes = [vp]
    lib.sdk_last_retry.restype = i64
    lib.sdk_merge_topk.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.sdk_stage_a_fetch.argtypes = [vp, vp, vp, i64, C.POINTER(i32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    _lib = lib
    return lib


def device_count() -> int:
    return load().sdk_device_count()


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().sdk_nccl_unique_id(buf)
    if rc != 0:
        raise NativeError(rc, load().sdk_last_error(None).decode())
    return buf.raw


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One matching context = one GPU (+ optionally one rank of a row-sharded bank)."""

    def __init__(self, device: int = 0, world: int = 1, rank: int = 0, nccl_uid: Optional[bytes] = None):
        self.lib = load()
        self.h = C.c_void_p()
        uid = C.create_string_buffer(nccl_uid, 128) if nccl_uid else None
        rc = self.lib.sdk_create(C.byref(self.h), device, world, rank, uid)
        if rc != 0:
            raise NativeError(rc, self.lib.sdk_last_error(None).decode())
        self.world, self.rank, self.device = world, rank, device
        self.L = self.k = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.sdk_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise NativeError(rc, self.lib.sdk_last_error(self.h).decode())

    def set_option(self, key: str, value: float):
        self._ck(self.lib.sdk_set_option(self.h, key.encode(), float(value)))

    # ---- bank ----
    def bank_load(self, rows, row_speaker, row_trust=None, dtype: int = DTYPE_F32, global_row_offset: int = 0):
        rows = _np(rows, np.float32)
        if rows.ndim != 2:
            raise ValueError("bank rows must be [P, D]")
        spk = _np(row_speaker, np.int32)
        tr = None if row_trust is None else _np(row_trust, np.uint8)
        if len(spk) != rows.shape[0] or (tr is not None and len(tr) != rows.shape[0]):
            raise ValueError("row_speaker / row_trust length must equal the number of bank rows")
        self._ck(self.lib.sdk_bank_load(self.h, _ptr(rows), _ptr(spk), _ptr(tr), rows.shape[0], rows.shape[1], dtype,
                                        global_row_offset))
        self.P, self.D = rows.shape

    def bank_load_dev(self, d_rows_ptr: int, d_spk_ptr: int, d_trust_ptr: Optional[int], P: int, D: int, dtype: int,
                      global_row_offset: int = 0):
        self._ck(self.lib.sdk_bank_load_dev(self.h, d_rows_ptr, d_spk_ptr, d_trust_ptr, P, D, dtype, global_row_offset))
        self.P, self.D = P, D

    # ---- identify ----
    def identify(self, seg, seg_label, L: int, pool: int = POOL_MEAN, threshold: float = 0.354, k: int = 10):
        """Host buffers in, host results out.  Returns (rows [L,k] int64, scores [L,k] f32, counts [L] i32).
        A float16 `seg` goes down as it is (sdk_identify_f16: half the host->device bytes, same result as on the
        widened matrix); anything else is taken as float32."""
        f16 = getattr(seg, "dtype", None) == np.float16
        seg = _np(seg, np.float16 if f16 else np.float32).reshape(-1, self.D)
        lab = _np(seg_label, np.int32)
        if len(lab) != seg.shape[0]:
            raise ValueError("seg_label length must equal the number of segments")
        rows = np.empty((L, k), dtype=np.int64)
        scores = np.empty((L, k), dtype=np.float32)
        counts = np.empty(L, dtype=np.int32)
        fn = self.lib.sdk_identify_f16 if f16 else self.lib.sdk_identify
        self._ck(fn(self.h, _ptr(seg), _ptr(lab), seg.shape[0], L, pool, threshold, k, _ptr(rows), _ptr(scores), _ptr(counts)))
        self.L, self.k = L, k
        return rows, scores, counts

    def identify_dev(self, d_seg_ptr: int, d_lab_ptr: int, N: int, L: int, pool: int, threshold: float, k: int, f16: bool = False):
        fn = self.lib.sdk_identify_f16_dev if f16 else self.lib.sdk_identify_dev
        self._ck(fn(self.h, d_seg_ptr, d_lab_ptr, N, L, pool, threshold, k))
        self.L, self.k = L, k

    def merge_topk(self, rows, scores, counts, trust=None):
        """Merge per-shard lists made elsewhere: rows [W,L,k] int64, scores [W,L,k] f32, counts [W,L] i32, trust
        [W,L,k] u8 or None.  The merged lists become this context's results (assign / fetch)."""
        rows = _np(rows, np.int64)
        W, L, k = rows.shape
        scores, counts = _np(scores, np.float32), _np(counts, np.int32)
        if scores.shape != (W, L, k) or counts.shape != (W, L):
            raise ValueError("scores must be [W,L,k] and counts [W,L]")
        tr = None if trust is None else _np(trust, np.uint8)
        self._ck(self.lib.sdk_merge_topk(self.h, W, L, k, _ptr(rows), _ptr(scores), _ptr(counts), _ptr(tr)))
        self.L, self.k = L, k

    def assign(self, assign_threshold: float = 0.3, min_trust: str = "low"):
        self._ck(self.lib.sdk_assign(self.h, assign_threshold, TRUST_CODES.get(min_trust, 99)))

    def fetch(self, with_assign: bool = False):
        L, k = self.L, self.k
        out = {"row": np.empty((L, k), np.int64), "score": np.empty((L, k), np.float32), "count": np.empty(L, np.int32),
               "trust": np.empty((L, k), np.uint8)}
        a = [None] * 5
        if with_assign:
            out.update({"assign_idx": np.empty(L, np.int32), "assign_score": np.empty(L, np.float64),
                        "assign_conf": np.empty(L, np.int32), "cand_idx": np.empty((L, 3), np.int32),
                        "cand_score": np.empty((L, 3), np.float64)})
            a = [out["assign_idx"], out["assign_score"], out["assign_conf"], out["cand_idx"], out["cand_score"]]
        self._ck(self.lib.sdk_results_fetch(self.h, _ptr(out["row"]), _ptr(out["score"]), _ptr(out["count"]),
                                            _ptr(out["trust"]), *[_ptr(x) for x in a]))
        return out

    # ---- config 5 ----
    def affinity_pooled(self, seg, seg_label, L: int, dtype: int = DTYPE_BF16, pool: int = POOL_MEAN, want_ll=True):
        seg = _np(seg, np.float32)
        lab = _np(seg_label, np.int32)
        N, D = seg.shape
        nl = np.empty((N, L), np.float32)
        ll = np.empty((L, L), np.float32) if want_ll else None
        self._ck(self.lib.sdk_affinity_pooled(self.h, _ptr(seg), _ptr(lab), N, D, L, dtype, pool, _ptr(nl), _ptr(ll)))
        return nl, ll

    def affinity_pooled_dev(self, d_seg_ptr, d_lab_ptr, N, D, L, dtype, pool, d_nl_ptr, d_ll_ptr=None):
        self._ck(self.lib.sdk_affinity_pooled_dev(self.h, d_seg_ptr, d_lab_ptr, N, D, L, dtype, pool, d_nl_ptr, d_ll_ptr))

    # ---- stream / timing ----
    def sync(self):
        self._ck(self.lib.sdk_sync(self.h))

    def stream_ptr(self) -> int:
        return int(self.lib.sdk_stream(self.h) or 0)

    def timer_start(self):
        self._ck(self.lib.sdk_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.sdk_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def profile_get(self, name: str):
        ms, n = C.c_float(), C.c_int64()
        self._ck(self.lib.sdk_profile_get(self.h, name.encode(), C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def profile_reset(self):
        self._ck(self.lib.sdk_profile_reset(self.h))

    def probe_bank_read(self):
        self._ck(self.lib.sdk_probe_bank_read(self.h))

    def launch_count(self) -> int:
        return int(self.lib.sdk_launch_count(self.h))

    def stage_a(self):
        """(rows [L,ncand] int32 shard-local (-1 pad), approx [L,ncand] f32, eps_base, eps_chain) of the last tensor-path identify."""
        n, eb, ec = C.c_int32(), C.c_float(), C.c_float()
        self._ck(self.lib.sdk_stage_a_fetch(self.h, None, None, 0, C.byref(n), C.byref(eb), C.byref(ec)))
        rows = np.empty((self.L, n.value), np.int32)
        approx = np.empty((self.L, n.value), np.float32)
        self._ck(self.lib.sdk_stage_a_fetch(self.h, _ptr(rows), _ptr(approx), rows.size, C.byref(n), C.byref(eb), C.byref(ec)))
        return rows, approx, float(eb.value), float(ec.value)

    def last_retry(self) -> int:
        return int(self.lib.sdk_last_retry(self.h))

    def last_path(self):
        p, f = C.c_int32(), C.c_int64()
        self._ck(self.lib.sdk_last_path(self.h, C.byref(p), C.byref(f)))
        return int(p.value), int(f.value)
