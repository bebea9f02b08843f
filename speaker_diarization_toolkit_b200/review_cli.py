"""`speaker-review consistency AUDIO` -- diarization consistency check on the pooled self-affinity (SURVEY 8f item 4,
BASELINE config 5).

The reference's `speaker-review` is an interactive TUI over transcript segments (speaker-review:948-966: review / status /
clear); it has no numeric check.  This adds ONE sub-command next to those: every segment's mean (or max) cosine affinity
to each diarization label of the same recording -- `sdk_affinity_pooled`, the [N, L] and [L, L] outputs of the tcgen05
accumulate-pooling kernel -- and reports the segments that sit closer to another label than to their own, i.e. the
ones a reviewer should listen to first.  Input: the recording's per-segment embedding sidecar (store.py).  There is no
CPU fallback: the affinity is computed on the GPU through the C-ABI.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np

from . import BACKEND_NAME, _native, store


def consistency(audio_path, backend_name: str = BACKEND_NAME, pool: str = "mean", margin: float = 0.0, dtype: str = "bf16",
                ctx: Optional[_native.Context] = None) -> Dict[str, Any]:
    """Returns {"labels", "label_affinity" [L][L], "segments_checked", "suspects": [...]}.
    A segment is a suspect when  affinity(own label) - max affinity(other label) < margin.  The segment itself takes part
    in its own label's pool (self-similarity 1/n_label of the mean), which only makes the check more conservative."""
    se = store.load_segment_embeddings(audio_path, backend_name)
    L, N = len(se.labels), int(se.emb.shape[0])
    own = ctx is None
    ctx = ctx or _native.Context(int(os.environ.get("SPEAKER_B200_DEVICE", 0)))
    try:
        nl, ll = ctx.affinity_pooled(se.emb, se.label_index, L, dtype=_native.DTYPE_BF16 if dtype == "bf16" else _native.DTYPE_F32,
                                     pool=_native.POOL_MAX if pool == "max" else _native.POOL_MEAN)
    finally:
        if own:
            ctx.close()
    suspects: List[Dict[str, Any]] = []
    if L > 1 and N:
        idx = np.arange(N)
        own_aff = nl[idx, se.label_index]
        others = nl.copy()
        others[idx, se.label_index] = -np.inf
        best_other = others.argmax(axis=1)
        best_aff = others[idx, best_other]
        gap = own_aff - best_aff
        for i in np.flatnonzero(gap < margin):
            suspects.append({"segment": int(i), "start": float(se.start[i]), "end": float(se.end[i]), "label": se.labels[int(se.label_index[i])],
                             "affinity": float(own_aff[i]), "closer_to": se.labels[int(best_other[i])],
                             "closer_affinity": float(best_aff[i]), "gap": float(gap[i])})
        suspects.sort(key=lambda r: (r["gap"], r["segment"]))
    return {"labels": se.labels, "label_affinity": [[float(v) for v in row] for row in ll], "segments_checked": N, "suspects": suspects,
            "pool": pool, "margin": margin, "backend": backend_name}


def cmd_consistency(args, ctx=None) -> int:
    audio = Path(args.audio)
    if not audio.exists():
        print(f"Error: Audio file not found: {audio}", file=sys.stderr)
        return 1
    backend_name = args.backend or os.environ.get("SPEAKER_DETECTION_BACKEND", BACKEND_NAME)
    try:
        rep = consistency(audio, backend_name, args.pool, args.margin, args.dtype, ctx)
    except Exception as exc:
        print(f"Error during consistency check: {exc}", file=sys.stderr)
        return 1
    if args.format == "json":
        print(json.dumps(rep, indent=2))
    else:
        print(f"{audio.name}: {rep['segments_checked']} segments, {len(rep['labels'])} labels, {len(rep['suspects'])} suspect segment(s)")
        for s in rep["suspects"][: args.limit]:
            print(f"  #{s['segment']} {s['start']:.2f}-{s['end']:.2f}s  {s['label']} ({s['affinity']:.3f}) closer to {s['closer_to']} "
                  f"({s['closer_affinity']:.3f})")
    return 2 if rep["suspects"] and args.fail_on_suspects else 0


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="speaker-review", description="diarization consistency check (B200 pooled affinity)")
    sub = parser.add_subparsers(dest="command")
    c = sub.add_parser("consistency", help="segments closer to another diarization label than to their own")
    c.add_argument("audio")
    c.add_argument("--backend", "-b")
    c.add_argument("--pool", choices=["mean", "max"], default="mean")
    c.add_argument("--margin", type=float, default=0.0, help="flag segments whose own-label affinity leads by less than this")
    c.add_argument("--dtype", choices=["bf16", "fp32"], default="bf16")
    c.add_argument("--format", "-f", choices=["text", "json"], default="text")
    c.add_argument("--limit", type=int, default=20)
    c.add_argument("--fail-on-suspects", action="store_true", help="exit code 2 when any segment is flagged")
    c.set_defaults(func=cmd_consistency)
    return parser


def main(argv=None) -> int:
    parser = build_parser()
    args = parser.parse_args(argv)
    if not args.command:
        parser.print_help()
        return 0
    return args.func(args)


if __name__ == "__main__":
    sys.exit(main())
