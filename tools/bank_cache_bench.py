#!/usr/bin/env python3
"""Wall time of the host-side bank load at scale (at a million profiles the reference's per-call
`list_all_speakers` + one .npy per record dwarfs the GPU time).

    python tools/bank_cache_bench.py --rows 1000000 --dim 512 [--dir /dev/shm/bankbench]

Builds a synthetic store (db/<id>.json + embeddings/<id>/<emb>.npy, one record per speaker), then times
  list_all_speakers          the reference's O(P) JSON pass (speaker_detection:206-220), restated in store.py
  list_all_speakers_cached   db/.profiles.pack: one scandir + stat pass, only changed files parsed again
  build_bank                 the uncached per-file load (what a backend without the pack would do)
  build_bank_cached  cold    first call: reads every .npy, writes the pack + binary index
  build_bank_cached  warm    per-file stat() + one dict pass + ONE gather from the mmap
  build_bank_cached  trust   SPEAKER_B200_BANK_CACHE=trust: no stat(), rows taken by key
Prints one JSON line.  CPU only; nothing here touches the GPU or the oracle."""
import argparse
import json
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--dir", default="/dev/shm/bankbench")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--skip-uncached", action="store_true")
    args = ap.parse_args()
    root = Path(args.dir)
    if root.exists():
        shutil.rmtree(root)
    (root / "db").mkdir(parents=True)
    os.environ["SPEAKERS_EMBEDDINGS_DIR"] = str(root)
    from speaker_diarization_toolkit_b200 import store
    rng = np.random.default_rng(1)
    P, D = args.rows, args.dim
    t0 = time.perf_counter()
    block = rng.standard_normal((min(P, 65536), D)).astype(np.float32)
    for i in range(P):
        sid = f"spk{i:07d}"
        d = root / "embeddings" / sid
        d.mkdir(parents=True)
        np.save(d / "e0.npy", block[i % len(block)])
        with open(root / "db" / f"{sid}.json", "w") as fh:
            json.dump({"id": sid, "name": sid, "tags": [], "embeddings": {"b200": [{"id": "e0", "trust_level": "high",
                                                                                     "model_version": "b200-canonical-v1"}]}}, fh)
    t_make = time.perf_counter() - t0
    res = {"rows": P, "dim": D, "store": str(root), "make_store_s": round(t_make, 2)}

    def timed(name, fn):
        t = time.perf_counter()
        out = fn()
        res[name] = round(time.perf_counter() - t, 3)
        return out

    speakers = timed("list_all_speakers_s", store.list_all_speakers)
    timed("list_all_speakers_cached_cold_s", store.list_all_speakers_cached)
    cached = timed("list_all_speakers_cached_warm_s", store.list_all_speakers_cached)
    res["profile_pack_identical"] = cached == speakers
    plain = None
    if not args.skip_uncached:
        plain = timed("build_bank_uncached_s", lambda: store.build_bank(speakers, "b200"))
    cold = timed("build_bank_cached_cold_s", lambda: store.build_bank_cached(speakers, "b200"))
    warm = timed("build_bank_cached_warm_s", lambda: store.build_bank_cached(speakers, "b200"))
    os.environ["SPEAKER_B200_BANK_CACHE"] = "trust"
    trust = timed("build_bank_cached_trust_s", lambda: store.build_bank_cached(speakers, "b200"))
    os.environ.pop("SPEAKER_B200_BANK_CACHE")
    t = time.perf_counter()
    touched = float(np.asarray(warm.rows).sum())            # pages the mmap in (what sdk_bank_load's H2D copy will read)
    res["touch_all_rows_s"] = round(time.perf_counter() - t, 3)
    ok = np.array_equal(cold.rows, warm.rows) and np.array_equal(warm.rows, trust.rows) and warm.speaker_ids == trust.speaker_ids
    if plain is not None:
        ok = ok and np.array_equal(plain.rows, warm.rows) and plain.speaker_ids == warm.speaker_ids
    res["identical"] = bool(ok)
    res["pack_bytes"] = (root / "embeddings" / f".bank-b200-D{D}.f32").stat().st_size
    res["index_bytes"] = (root / "embeddings" / f".bank-b200-D{D}.idx.npz").stat().st_size
    res["host_cores"] = os.cpu_count()
    print(json.dumps(res))
    if not args.keep:
        shutil.rmtree(root)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
