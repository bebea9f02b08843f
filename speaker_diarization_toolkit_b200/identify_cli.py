"""`speaker_detection identify|verify` for the b200 backend -- the caller of the hot path.

Mirrors cmd_identify (speaker_detection:1031-1133) and cmd_verify (:1136-1178): same flags (:1497-1513), same
stderr strings and return codes, stdout = pure JSON with the same keys.  One addition: rows a backend tags with a
diarization `label` keep that key (and `rank`) in the JSON, which is how per-label results reach `speaker-assign`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path
from typing import Any, Dict, List, Optional

from . import store
from .plugin_api import get_backend

_TRUST_RANK = {"high": 3, "medium": 2, "low": 1, "unknown": 0, "invalidated": -1}


def default_backend_name(cli_value: Optional[str] = None) -> str:
    return cli_value or os.environ.get("SPEAKER_DETECTION_BACKEND", "b200")


def decorate_results(results: List[Dict[str, Any]], backend_name: str) -> List[Dict[str, Any]]:
    """speaker_detection:1082-1123: add name / trust_level / embedding_id from the stored profile."""
    out = []
    for r in results:
        sid = r["speaker_id"]
        profile = store.load_speaker(sid)
        conf = r.get("confidence", r.get("similarity", 0))
        emb_id = r.get("embedding_id")
        trust = "unknown"
        if profile:
            records = profile.get("embeddings", {}).get(backend_name, [])
            if emb_id:
                for rec in records:
                    if rec.get("id") == emb_id:
                        trust = rec.get("trust_level", "unknown")
                        break
            elif records:
                best, best_id = "unknown", None
                for rec in records:
                    t = rec.get("trust_level", "unknown")
                    if _TRUST_RANK.get(t, 0) > _TRUST_RANK.get(best, 0):
                        best, best_id = t, rec.get("id")
                trust, emb_id = best, best_id
        row = {
            "speaker_id": sid,
            "name": profile["names"]["default"] if profile else sid,
            "score": conf,
            "confidence": conf,
            "trust_level": trust,
            "embedding_id": emb_id,
            "backend": backend_name,
        }
        if r.get("label") is not None:
            row["label"] = r["label"]
            if "rank" in r:
                row["rank"] = r["rank"]
        out.append(row)
    return out


def identify_rows(audio_path: Path, backend_name: str, tags: Optional[str], threshold: float, backend=None):
    """Shared by the CLI and by speaker-assign's in-process embedding step.
    Returns (rc, rows, message): rc 0 with rows, or rc 1 with the reference's stderr message."""
    speakers = store.list_all_speakers()
    if tags:
        speakers = store.filter_speakers_by_tags(speakers, [t.strip() for t in tags.split(",")], any_tag=False)
    if not speakers:
        return 1, [], "No speakers to match against."
    candidates = [s for s in speakers if s.get("embeddings", {}).get(backend_name)]
    if not candidates:
        return 1, [], f"No speakers with {backend_name} embeddings."
    if backend is None:
        try:
            backend = get_backend(backend_name)
        except (ValueError, ImportError) as exc:
            return 1, [], f"Error loading backend: {exc}"
    print(f"Identifying speaker in {audio_path.name} against {len(candidates)} candidates...", file=sys.stderr)
    try:
        results = backend.identify_speaker(audio_path, candidates, threshold)
    except Exception as exc:
        return 1, [], f"Error during identification: {exc}"
    return 0, decorate_results(results, backend_name), ""


def cmd_identify(args, backend=None) -> int:
    audio_path = Path(args.audio)
    if not audio_path.exists():
        print(f"Error: Audio file not found: {audio_path}", file=sys.stderr)
        return 1
    backend_name = default_backend_name(args.backend)
    rc, rows, msg = identify_rows(audio_path, backend_name, args.tags, args.threshold, backend)
    if rc != 0:
        print(msg, file=sys.stderr)
        return rc
    as_json = getattr(args, "format", "text") == "json"
    if not rows:
        print("[]" if as_json else "No matching speakers found.")
        return 0
    if as_json:
        print(json.dumps(rows, indent=2))
    else:
        print("\nMatches:")
        for item in rows:
            tag = f" [{item['label']}]" if "label" in item else ""
            print(f"  {item['speaker_id']}: {item['name']} (confidence: {item['score']:.2f}){tag}")
    return 0


def cmd_verify(args, backend=None) -> int:
    speaker_id = args.id.lower().replace(" ", "-")
    profile = store.load_speaker(speaker_id)
    if not profile:
        print(f"Error: Speaker '{speaker_id}' not found.", file=sys.stderr)
        return 1
    audio_path = Path(args.audio)
    if not audio_path.exists():
        print(f"Error: Audio file not found: {audio_path}", file=sys.stderr)
        return 1
    backend_name = default_backend_name(args.backend)
    if not profile.get("embeddings", {}).get(backend_name):
        print(f"Error: Speaker '{speaker_id}' has no {backend_name} embeddings.", file=sys.stderr)
        return 1
    if backend is None:
        try:
            backend = get_backend(backend_name)
        except (ValueError, ImportError) as exc:
            print(f"Error loading backend: {exc}", file=sys.stderr)
            return 1
    print(f"Verifying audio against speaker '{speaker_id}'...", file=sys.stderr)
    try:
        result = backend.verify_speaker(audio_path, profile, args.threshold)
    except Exception as exc:
        print(f"Error during verification: {exc}", file=sys.stderr)
        return 1
    if result["match"]:
        print(f"MATCH: Speaker '{speaker_id}' verified (confidence: {result['confidence']:.2f})")
        return 0
    print(f"NO MATCH: Audio does not match speaker '{speaker_id}'")
    return 1


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="speaker_detection", description="identify / verify on the B200 matching path")
    sub = parser.add_subparsers(dest="command")
    p = sub.add_parser("identify", help="Identify speaker in audio")
    p.add_argument("audio")
    p.add_argument("--backend", "-b")
    p.add_argument("--tags")
    p.add_argument("--threshold", type=float, default=0.354)
    p.add_argument("--format", "-f", choices=["text", "json"], default="text")
    p.set_defaults(func=cmd_identify)
    v = sub.add_parser("verify", help="Verify speaker in audio")
    v.add_argument("id")
    v.add_argument("audio")
    v.add_argument("--backend", "-b")
    v.add_argument("--threshold", type=float, default=0.354)
    v.set_defaults(func=cmd_verify)
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
