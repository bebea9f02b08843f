"""`speaker_detection identify|verify|enroll` for the b200 backend -- the caller of the hot path and
the bank-writing path next to it.

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


def identify_rows(audio_path: Path, backend_name: str, tags: Optional[str], threshold: float, backend=None, status: bool = True):
    """Shared by the CLI and by speaker-assign's in-process embedding step (status=False there: the reference's assign
    runs identify as a subprocess and swallows its stderr, speaker-assign:288-293).
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
    if status:
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


def parse_segments(segments_str: str):
    """'10.5:45.2,78:120' -> [(10.5, 45.2), (78.0, 120.0)]  (speaker_detection:731-751, same messages)."""
    segments = []
    for part in segments_str.split(","):
        part = part.strip()
        if ":" not in part:
            raise ValueError(f"Invalid segment format '{part}'. Use 'start:end'.")
        a, b = part.split(":", 1)
        try:
            start, end = float(a), float(b)
        except ValueError:
            raise ValueError(f"Invalid segment times '{part}'. Must be numeric.")
        if start >= end:
            raise ValueError(f"Invalid segment '{part}'. Start must be < end.")
        segments.append((start, end))
    return segments


def cmd_enroll(args, backend=None) -> int:
    """Mirror of cmd_enroll (speaker_detection:754-919): same flags, messages, return codes and embedding record.
    Two things the reference CLI cannot do for a local backend are done here: the vector is also written
    to the canonical embeddings/<speaker_id>/<emb_id>.npy, and a vector whose dimension differs from the bank already
    enrolled for the backend is rejected before anything is stored."""
    import uuid
    from datetime import datetime, timezone
    import numpy as np
    speaker_id = store.normalize_speaker_id(args.id)
    profile = store.load_speaker(speaker_id)
    if not profile:
        print(f"Error: Speaker '{speaker_id}' not found. Use 'add' first.", file=sys.stderr)
        return 1
    audio_path = Path(args.audio)
    if not audio_path.exists():
        print(f"Error: Audio file not found: {audio_path}", file=sys.stderr)
        return 1
    backend_name = default_backend_name(args.backend)
    segments = None
    if args.segments:
        try:
            segments = parse_segments(args.segments)
        except ValueError as exc:
            print(f"Error: {exc}", file=sys.stderr)
            return 1
    elif args.from_transcript:
        tpath = Path(args.from_transcript)
        if not tpath.exists():
            print(f"Error: Transcript file not found: {tpath}", file=sys.stderr)
            return 1
        if not args.speaker_label:
            print("Error: --speaker-label required with --from-transcript", file=sys.stderr)
            return 1
        try:
            from . import transcript as _t
            segments = _t.extract_segments_as_tuples(_t.load_transcript(tpath), args.speaker_label)
        except Exception as exc:
            print(f"Error extracting segments: {exc}", file=sys.stderr)
            return 1
        if not segments:
            print(f"Error: No segments found for speaker '{args.speaker_label}' in transcript.", file=sys.stderr)
            return 1
        total = sum(e - s for s, e in segments)
        print(f"Found {len(segments)} segments for speaker '{args.speaker_label}' ({total:.1f}s total)", file=sys.stderr)
    elif getattr(args, "from_stdin", False):
        segments = []
        try:
            for line in sys.stdin:
                line = line.strip()
                if not line:
                    continue
                rec = json.loads(line)
                if rec.get("start") is not None and rec.get("end") is not None:
                    segments.append((float(rec["start"]), float(rec["end"])))
        except json.JSONDecodeError as exc:
            print(f"Error parsing JSONL from stdin: {exc}", file=sys.stderr)
            return 1
        if not segments:
            print("Error: No segments read from stdin. Provide JSONL with 'start' and 'end' fields.", file=sys.stderr)
            return 1
        total = sum(e - s for s, e in segments)
        print(f"Read {len(segments)} segments from stdin ({total:.1f}s total)", file=sys.stderr)
    if getattr(args, "dry_run", False):
        print(f"Would enroll speaker: {speaker_id}")
        print(f"  Audio: {audio_path}")
        print(f"  Backend: {backend_name}")
        if segments:
            total = sum(e - s for s, e in segments)
            print(f"  Segments: {len(segments)} ({total:.1f}s total)")
            for i, (a, b) in enumerate(segments[:5]):
                print(f"    {i+1}. {a:.2f}s - {b:.2f}s ({b-a:.2f}s)")
            if len(segments) > 5:
                print(f"    ... and {len(segments) - 5} more")
        return 0
    if backend is None:
        try:
            backend = get_backend(backend_name)
        except ValueError as exc:
            print(f"Error: {exc}", file=sys.stderr)
            return 1
        except ImportError as exc:
            print(f"Error loading backend '{backend_name}': {exc}", file=sys.stderr)
            return 1
    if not getattr(args, "quiet", False):
        print(f"Enrolling speaker '{speaker_id}' using {backend_name}...", file=sys.stderr)
    try:
        result = backend.enroll_speaker(audio_path, segments)
        vec = None
        if result.get("file") and Path(result["file"]).exists():
            vec = np.load(result["file"]).astype(np.float32).reshape(-1)
            have = store.bank_dimension(backend_name)
            if have is not None and have != vec.shape[0]:
                raise ValueError(f"embedding is {vec.shape[0]}-d but the bank enrolled for {backend_name} is {have}-d")
    except Exception as exc:
        print(f"Error during enrollment: {exc}", file=sys.stderr)
        return 1
    emb_id = f"emb-{uuid.uuid4().hex[:8]}"
    from .assign_cli import compute_b3sum
    audio_b3sum = compute_b3sum(audio_path)
    samples = store.get_samples_by_source_audio(speaker_id, audio_b3sum)
    trust_level = getattr(args, "trust_level", None) or store.compute_trust_level(samples)
    record = {
        "id": emb_id,
        "external_id": result.get("external_id"),
        "source_audio": str(audio_path.resolve()),
        "source_audio_b3sum": audio_b3sum,
        "source_segments": [{"start": s, "end": e} for s, e in segments] if segments else None,
        "model_version": result.get("model_version", "unknown"),
        "samples": samples,
        "trust_level": trust_level,
        "created_at": datetime.now(timezone.utc).isoformat(),
    }
    if "all_identifiers" in result:
        record["all_identifiers"] = result["all_identifiers"]
    if vec is not None:
        store.store_vector_canonical(speaker_id, emb_id, vec)
    profile.setdefault("embeddings", {}).setdefault(backend_name, []).append(record)
    store.save_speaker(profile)
    tracked = len(samples["reviewed"]) + len(samples["unreviewed"])
    if tracked > 0:
        print(f"Enrolled embedding {emb_id} for speaker '{speaker_id}' (trust: {trust_level}, {tracked} samples tracked)")
    else:
        print(f"Enrolled embedding {emb_id} for speaker '{speaker_id}' (no samples tracked)")
    return 0


def cmd_validate(args) -> int:
    """Mirror of cmd_validate (speaker_detection:1307-1361): schema check of the stored profiles and their embedding
    records with this package's validators (schemas.py), same output and return codes."""
    from . import schemas
    if args.speaker_id:
        speaker_id = store.normalize_speaker_id(args.speaker_id)
        profile = store.load_speaker(speaker_id)
        if not profile:
            print(f"Error: Speaker '{speaker_id}' not found.", file=sys.stderr)
            return 1
        speakers = [profile]
    else:
        speakers = store.list_all_speakers()
    if not speakers:
        print("No speakers found.")
        return 0
    total_warnings = profiles_with_issues = 0
    for profile in speakers:
        speaker_id = profile["id"]
        warnings = schemas.validate_profile(profile, strict=False)
        if warnings:
            profiles_with_issues += 1
            total_warnings += len(warnings)
            if args.verbose or not args.quiet:
                print(f"\n{speaker_id}:")
                for w in warnings:
                    print(f"  - {w}")
        elif args.verbose:
            print(f"{speaker_id}: OK")
    if not args.quiet:
        print(f"\nValidated {len(speakers)} profiles")
        if total_warnings > 0:
            print(f"  {profiles_with_issues} profiles with issues")
            print(f"  {total_warnings} total warnings")
        else:
            print("  All profiles valid")
    return 1 if total_warnings > 0 and args.strict else 0


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="speaker_detection", description="identify / verify / enroll / validate on the B200 matching path")
    parser.add_argument("-q", "--quiet", action="store_true", help="Suppress status messages")   # speaker_detection:1374
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
    e = sub.add_parser("enroll", help="Enroll speaker from audio")       # flags of speaker_detection:1456-1473
    e.add_argument("id", help="Speaker ID")
    e.add_argument("audio", help="Audio file path")
    e.add_argument("--backend", "-b")
    e.add_argument("--segments", "-s", help="Time segments 'start:end,start:end' (seconds)")
    e.add_argument("--from-transcript", "-t", metavar="JSON")
    e.add_argument("--speaker-label", "-l")
    e.add_argument("--from-stdin", action="store_true")
    e.add_argument("-n", "--dry-run", action="store_true")
    e.add_argument("--trust-level", choices=["high", "medium", "low"])
    e.set_defaults(func=cmd_enroll)
    c = sub.add_parser("validate", help="Validate schema of profiles and embeddings")      # speaker_detection:1525-1534
    c.add_argument("speaker_id", nargs="?", help="Speaker ID (optional, validates all if omitted)")
    c.add_argument("-v", "--verbose", action="store_true", help="Show OK profiles too")
    c.add_argument("-q", "--quiet", action="store_true", help="Only show summary")
    c.add_argument("--strict", action="store_true", help="Return non-zero exit code on warnings")
    c.set_defaults(func=cmd_validate)
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
