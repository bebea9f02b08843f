"""`speaker-assign assign` with a per-label embedding step (option (ii) of 8b).

Same flags (speaker-assign:747-760), same YAML/JSON output (:598-649) and the same `combine_signals`
semantics as the reference.  Differences, both on the embedding step only:
  * ONE identify call per recording, in-process, instead of one `speaker_detection identify` subprocess per
    label (speaker-assign:283-294);
  * a label receives only the rows whose `label` equals it (the reference hands every label the same
    whole-recording list, speaker-assign:276-278 "simplified implementation").
The LLM signal (`--use-llm`) stays a subprocess to `speaker-llm` exactly as in the reference (:356-400); it is
outside the hot path.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
from datetime import datetime, timezone
from pathlib import Path
from typing import List, Optional

from . import signals as sg
from . import store, transcript
from .identify_cli import default_backend_name, identify_rows

VERSION = "1.0.0"
SCHEMA_VERSION = 1

try:
    import yaml
    _YAML = True
except ImportError:  # pragma: no cover
    _YAML = False


def compute_b3sum(path: Path) -> str:
    """Blake3 via the `b3sum` binary, SHA-256 fallback, first 32 hex digits (speaker-assign:102-118)."""
    try:
        r = subprocess.run(["b3sum", "--no-names", str(path)], capture_output=True, text=True, check=True)
        return r.stdout.strip()[:32]
    except (subprocess.CalledProcessError, FileNotFoundError):
        h = hashlib.sha256()
        with open(path, "rb") as fh:
            for chunk in iter(lambda: fh.read(1 << 16), b""):
                h.update(chunk)
        return h.hexdigest()[:32]


def _load_yaml(path: Path) -> dict:
    text = path.read_text()
    return (yaml.safe_load(text) or {}) if _YAML else json.loads(text)


def _save_yaml(path: Path, data: dict) -> None:
    with open(path, "w") as fh:
        if _YAML:
            yaml.dump(data, fh, default_flow_style=False, sort_keys=False, allow_unicode=True)
        else:
            json.dump(data, fh, indent=2, ensure_ascii=False)


def collect_llm_signals(speaker_label: str, transcript_path: Path, context_name: Optional[str]) -> List[sg.Signal]:
    out: List[sg.Signal] = []
    cmd = ["speaker-llm", "analyze", str(transcript_path)] + (["--context", context_name] if context_name else [])
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, env=os.environ.copy())
    except FileNotFoundError:
        return out
    if r.returncode != 0 or not r.stdout.strip():
        return out
    try:
        analysis = json.loads(r.stdout)
    except json.JSONDecodeError:
        return out
    for det in analysis.get("detections", []):
        if det.get("speaker_label") == speaker_label:
            out.append(sg.Signal("llm_name_detection", det.get("detected_name", "").lower().replace(" ", "-"),
                                 det.get("confidence", 0.5),
                                 {"detected_name": det.get("detected_name"), "evidence": det.get("evidence", [])}))
    return out


def assign_labels(audio_path: Path, transcript_data: dict, *, use_embeddings: bool, min_trust: str, tags: Optional[str],
                  threshold: float, expected_speakers: List[str], context_name: Optional[str], use_llm: bool,
                  transcript_path: Optional[Path] = None, verbose: bool = False, backend=None) -> dict:
    """The label loop of cmd_assign (speaker-assign:543-595) -> `mappings` dict."""
    labels = transcript.get_speakers_from_transcript(transcript_data)
    rows = None
    if use_embeddings:
        # identify is never given assign's --threshold / --backend by the reference either:
        # backend from $SPEAKER_DETECTION_BACKEND, threshold 0.354
        rc, rows, msg = identify_rows(audio_path, default_backend_name(None), tags, 0.354, backend, status=False)
        if rc != 0:
            if verbose:
                print(f"  identify: {msg}", file=sys.stderr)
            rows = None       # graceful degradation, as speaker-assign:296-326
    mappings = {}
    for label in labels:
        segs = transcript.get_speaker_segments(transcript_data, label)
        if verbose:
            print(f"\nProcessing speaker {label} ({len(segs)} segments)...")
        sigs: List[sg.Signal] = []
        if use_embeddings:                                  # (the progress lines of speaker-assign:558-590, verbatim)
            if verbose:
                print("  Collecting embedding signals...")
            emb = sg.signals_from_matches(rows, min_trust=min_trust, label=label) if rows is not None else []
            sigs.extend(emb)
            if verbose:
                for s in emb:
                    print(f"    - {s.speaker_id}: {s.score:.2f} (trust: {s.evidence.get('trust_level', '-')})")
        if expected_speakers:
            if verbose:
                print("  Collecting context signals...")
            sigs.extend(sg.collect_context_signals(label, context_name, expected_speakers))
        if use_llm and transcript_path is not None:
            if verbose:
                print("  Collecting LLM signals...")
            llm = collect_llm_signals(label, transcript_path, context_name)
            sigs.extend(llm)
            if verbose:
                for s in llm:
                    print(f"    - {s.speaker_id}: {s.score:.2f}")
        a = sg.combine_signals(label, sigs, threshold=threshold)
        mappings[label] = {"speaker_id": a.speaker_id, "confidence": a.confidence, "score": round(a.score, 3),
                           "signals": a.signals}
        if a.candidates:
            mappings[label]["candidates"] = a.candidates
    return mappings


def cmd_assign(args, backend=None) -> int:
    audio_path = Path(args.audio).resolve()
    transcript_path = Path(args.transcript).resolve()
    if not audio_path.exists():
        print(f"Error: Audio file not found: {audio_path}", file=sys.stderr)
        return 1
    if not transcript_path.exists():
        print(f"Error: Transcript file not found: {transcript_path}", file=sys.stderr)
        return 1
    with open(transcript_path, "r") as fh:
        data = json.load(fh)
    labels = transcript.get_speakers_from_transcript(data)
    if not labels:
        print("Error: No speakers found in transcript", file=sys.stderr)
        return 1
    if not args.quiet:
        print(f"Found {len(labels)} speakers: {', '.join(labels)}")
    context_name, expected = args.context, []
    b3 = compute_b3sum(audio_path)
    cat = store.get_db_dir() / "catalog" / f"{b3}.yaml"
    if cat.exists():
        entry = _load_yaml(cat)
        context_name = context_name or entry.get("context", {}).get("name")
        expected = entry.get("context", {}).get("expected_speakers", [])
    if args.expected_speakers:
        expected = args.expected_speakers.split(",")
    mappings = assign_labels(audio_path, data, use_embeddings=args.use_embeddings, min_trust=args.min_trust, tags=args.tags,
                             threshold=args.threshold, expected_speakers=expected, context_name=context_name,
                             use_llm=args.use_llm, transcript_path=transcript_path, verbose=args.verbose, backend=backend)
    output = {
        "schema_version": SCHEMA_VERSION,
        "recording_b3sum": b3,
        "transcript_path": str(transcript_path),
        "assigned_at": datetime.now(timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ"),
        "method": f"speaker-assign-v{VERSION}",
        "context": context_name,
        "min_trust": args.min_trust,
        "threshold": args.threshold,
        "mappings": mappings,
    }

    def text_lines():
        for label, d in mappings.items():
            yield f"  {label} -> {d.get('speaker_id') or '(unassigned)'} ({d.get('confidence', '?')}, score: {d.get('score', 0):.2f})"

    if args.dry_run:
        print("\n=== DRY RUN - No changes saved ===")
        if args.format == "json":
            print(json.dumps(output, indent=2, ensure_ascii=False))
        else:
            print(f"\nAssignments for: {audio_path.name}")
            print("-" * 50)
            for label, d in mappings.items():
                print(f"  {label} -> {d.get('speaker_id') or '(unassigned)'} ({d.get('confidence', '?')}, score: {d.get('score', 0):.2f})")
                if d.get("candidates"):
                    print(f"       candidates: {', '.join(c['speaker_id'] for c in d['candidates'])}")
        return 0
    adir = store.get_db_dir() / "assignments"
    adir.mkdir(parents=True, exist_ok=True)
    apath = adir / f"{b3}.yaml"
    _save_yaml(apath, output)
    if args.output:
        _save_yaml(Path(args.output), output)
    if args.format == "json":
        print(json.dumps(output, indent=2, ensure_ascii=False))
    elif not args.quiet:
        print(f"\nAssignments saved: {apath.name}")
        print("-" * 50)
        print(f"Assigned: {sum(1 for a in mappings.values() if a.get('speaker_id'))}/{len(mappings)}")
        for line in text_lines():
            print(line)
    return 0


def cmd_assign_batch(args, matcher=None) -> int:
    """Many recordings, one resident backend call (replaces one `speaker-assign assign` process tree
    per recording, speaker-process:478-509).  MANIFEST: JSON list or JSON-lines of {"audio": ..., "transcript": ...}.
    Writes the same assignments/<b3sum>.yaml per recording as `assign --use-embeddings`."""
    from .batch import BatchMatcher
    text = Path(args.manifest).read_text()
    try:
        items = json.loads(text)
    except json.JSONDecodeError:
        items = [json.loads(line) for line in text.splitlines() if line.strip()]
    if getattr(args, "gpus", 1) > 1 and matcher is None:
        return _assign_batch_multi_gpu(args, items)
    audios, transcripts, b3s, expected = [], [], [], []
    for it in items:
        a, t = Path(it["audio"]).resolve(), Path(it["transcript"]).resolve()
        if not a.exists():
            print(f"Error: Audio file not found: {a}", file=sys.stderr)
            return 1
        if not t.exists():
            print(f"Error: Transcript file not found: {t}", file=sys.stderr)
            return 1
        b3 = compute_b3sum(a)
        ctx_name, exp = None, []
        cat = store.get_db_dir() / "catalog" / f"{b3}.yaml"
        if cat.exists():
            entry = _load_yaml(cat)
            ctx_name = entry.get("context", {}).get("name")
            exp = entry.get("context", {}).get("expected_speakers", [])
        audios.append(a); transcripts.append(t); b3s.append(b3); expected.append((ctx_name, exp) if exp else None)
    own = matcher is None
    matcher = matcher or BatchMatcher()
    try:
        matcher.load_bank(args.tags)
        results = matcher.identify(audios, assign_threshold=args.threshold, min_trust=args.min_trust, expected=expected)
    except Exception as exc:
        print(f"Error during identification: {exc}", file=sys.stderr)
        return 1
    finally:
        if own:
            matcher.close()
    outputs = []
    for a, t, b3, res, exp in zip(audios, transcripts, b3s, results, expected):
        with open(t, "r") as fh:
            labels = transcript.get_speakers_from_transcript(json.load(fh))
        mappings = {}
        for label in labels:
            m = res.mappings.get(label)
            if m is None:       # a transcript label without segment embeddings: no signal
                m = {"speaker_id": None, "confidence": "unassigned", "score": 0.0, "signals": []}
            mappings[label] = m
        out = {"schema_version": SCHEMA_VERSION, "recording_b3sum": b3, "transcript_path": str(t),
               "assigned_at": datetime.now(timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ"), "method": f"speaker-assign-v{VERSION}",
               "context": exp[0] if exp else None, "min_trust": args.min_trust, "threshold": args.threshold, "mappings": mappings}
        outputs.append(out)
        if not args.dry_run:
            adir = store.get_db_dir() / "assignments"
            adir.mkdir(parents=True, exist_ok=True)
            _save_yaml(adir / f"{b3}.yaml", out)
    if args.format == "json":
        print(json.dumps(outputs, indent=2, ensure_ascii=False))
    elif not args.quiet:
        for a, out in zip(audios, outputs):
            done = sum(1 for m in out["mappings"].values() if m.get("speaker_id"))
            print(f"{a.name}: assigned {done}/{len(out['mappings'])}")
    return 0


def _assign_batch_multi_gpu(args, items) -> int:
    """`assign-batch --gpus N`: one worker process per GPU (replaces the thread pool of per-recording process
    trees at speaker-process:627-642).  Default: recordings are split over the GPUs, balanced by segment count, every
    worker holds the whole bank -- no collective.  `--shard-bank`: every worker loads its slice of the bank rows and
    scores all recordings; the library all-gathers the per-shard top-k (NCCL) and merges, rank 0 writes the output."""
    import subprocess
    import tempfile
    from . import BACKEND_NAME, sharding
    n = int(args.gpus)
    from . import _native
    n_dev = max(1, _native.device_count())       # more workers than GPUs: they share devices (data-parallel mode only)
    if args.shard_bank and n > n_dev:
        print(f"Error: --shard-bank needs one GPU per rank ({n} requested, {n_dev} visible)", file=sys.stderr)
        return 1
    pkg_root = str(Path(__file__).resolve().parent.parent)          # the workers import this package by name
    os.environ["PYTHONPATH"] = pkg_root + (os.pathsep + os.environ["PYTHONPATH"] if os.environ.get("PYTHONPATH") else "")
    base = [sys.executable, "-m", "speaker_diarization_toolkit_b200.assign_cli", "assign-batch"]
    common = ["--gpus", "1", "--min-trust", args.min_trust, "--threshold", repr(float(args.threshold)), "--format", "json"]
    if args.tags:
        common += ["--tags", args.tags]
    with tempfile.TemporaryDirectory(prefix="assign-batch-") as td:
        jobs = []
        if args.shard_bank:
            man = Path(td) / "manifest.json"
            man.write_text(json.dumps(items))
            for r in range(n):
                env = dict(os.environ, SPEAKER_B200_DEVICE=str(r % n_dev), SPEAKER_B200_WORLD=str(n), SPEAKER_B200_RANK=str(r),
                           SPEAKER_B200_UID_FILE=str(Path(td) / "nccl.uid"))
                extra = ["--dry-run"] if (args.dry_run or r != 0) else []
                jobs.append((r, env, base + [str(man)] + common + extra))
        else:
            try:
                counts = [store.sidecar_segment_count(Path(it["audio"]).resolve(), BACKEND_NAME) for it in items]
            except (OSError, KeyError, ValueError) as exc:
                print(f"Error during identification: {exc}", file=sys.stderr)
                return 1
            for r, (a, b) in enumerate(sharding.partition_recordings(counts, n)):
                if a == b:
                    continue
                man = Path(td) / f"manifest.{r}.json"
                man.write_text(json.dumps(items[a:b]))
                env = dict(os.environ, SPEAKER_B200_DEVICE=str(r % n_dev))
                for v in ("SPEAKER_B200_WORLD", "SPEAKER_B200_RANK", "SPEAKER_B200_UID_FILE"):
                    env.pop(v, None)
                jobs.append((r, env, base + [str(man)] + common + (["--dry-run"] if args.dry_run else [])))
        procs = [(r, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)) for r, env, cmd in jobs]
        outputs, rc = [], 0
        for r, pr in procs:
            out, err = pr.communicate()
            if pr.returncode != 0:
                print(f"Error: GPU worker {r} failed (exit {pr.returncode}):\n{err.strip()}", file=sys.stderr)
                rc = 1
                continue
            if err.strip() and args.verbose:
                print(err.strip(), file=sys.stderr)
            if not args.shard_bank or r == 0:
                # (libraries may write to the worker's stdout as well -- NCCL prints its version line there when NCCL_DEBUG
                #  is set: take the JSON list the worker printed, not the whole stream)
                a = out.find("[\n") if "[\n" in out else out.find("[]")
                b = out.rfind("]")
                if a < 0 or b < a:
                    print(f"Error: GPU worker {r} printed no result list:\n{out.strip()[:400]}\n{err.strip()[:400]}", file=sys.stderr)
                    rc = 1
                    continue
                outputs += json.loads(out[a:b + 1])
    if rc:
        return rc
    if args.format == "json":
        print(json.dumps(outputs, indent=2, ensure_ascii=False))
    elif not args.quiet:
        for it, out in zip(items, outputs):
            done = sum(1 for m in out["mappings"].values() if m.get("speaker_id"))
            print(f"{Path(it['audio']).name}: assigned {done}/{len(out['mappings'])}")
    return 0


def resolve_audio_b3sum(audio_arg: str) -> Optional[str]:
    """A catalogued b3sum prefix (6..32 hex digits, unique) or a path to hash (speaker-assign:145-162)."""
    if 6 <= len(audio_arg) <= 32 and all(ch in "0123456789abcdef" for ch in audio_arg.lower()):
        matches = list((store.get_db_dir() / "catalog").glob(f"{audio_arg.lower()}*.yaml"))
        if len(matches) == 1:
            return matches[0].stem
        if len(matches) > 1:
            print(f"Error: Ambiguous b3sum prefix '{audio_arg}'", file=sys.stderr)
            return None
    path = Path(audio_arg).resolve()
    return compute_b3sum(path) if path.exists() else None


def cmd_show(args) -> int:
    """Mirror of cmd_show (speaker-assign:652-702): the stored assignments of a recording as text / JSON / YAML."""
    b3 = resolve_audio_b3sum(args.audio)
    if not b3:
        print(f"Error: Could not resolve audio: {args.audio}", file=sys.stderr)
        return 1
    path = store.get_db_dir() / "assignments" / f"{b3}.yaml"
    if not path.exists():
        print("Error: No assignments found for this recording", file=sys.stderr)
        return 1
    data = _load_yaml(path)
    if args.format == "json" or (args.format == "yaml" and not _YAML):
        print(json.dumps(data, indent=2, ensure_ascii=False))
    elif args.format == "yaml":
        print(yaml.dump(data, default_flow_style=False, sort_keys=False, allow_unicode=True))
    else:
        print(f"Assignments for: {b3[:8]}...")
        print(f"Context: {data.get('context') or '-'}")
        print(f"Method: {data.get('method', '-')}")
        print(f"Assigned at: {data.get('assigned_at', '-')}")
        print(f"Threshold: {data.get('threshold', '-')}")
        print(f"Min trust: {data.get('min_trust', '-')}")
        print()
        mappings = data.get("mappings", {})
        if not mappings:
            print("No mappings found")
        else:
            print("Mappings:")
            for label, info in mappings.items():
                print(f"  {label} -> {info.get('speaker_id') or '(unassigned)'}")
                print(f"       confidence: {info.get('confidence', '?')}, score: {info.get('score', 0):.3f}")
                if info.get("signals"):
                    print(f"       signals: {len(info['signals'])}")
                    for sig in info["signals"][:3]:
                        print(f"         - {sig.get('type', '?')}: {sig.get('score', 0):.2f}")
                if info.get("candidates"):
                    cands = ", ".join(f"{c['speaker_id']}({c['score']:.2f})" for c in info["candidates"])
                    print(f"       candidates: {cands}")
    return 0


def cmd_clear(args) -> int:
    """Mirror of cmd_clear (speaker-assign:705-728)."""
    b3 = resolve_audio_b3sum(args.audio)
    if not b3:
        print(f"Error: Could not resolve audio: {args.audio}", file=sys.stderr)
        return 1
    path = store.get_db_dir() / "assignments" / f"{b3}.yaml"
    if not path.exists():
        print("No assignments found for this recording", file=sys.stderr)
        return 0
    if not args.force:
        print(f"Clear assignments for: {b3[:8]}...?")
        if input("Confirm [y/N]: ").lower() != "y":
            print("Cancelled")
            return 0
    path.unlink()
    if not args.quiet:
        print(f"Cleared assignments: {b3[:8]}...")
    return 0


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="speaker-assign", description="Multi-signal speaker name assignment (B200 embedding path)")
    parser.add_argument("-V", "--version", action="version", version=f"speaker-assign {VERSION}")
    parser.add_argument("-q", "--quiet", action="store_true")
    parser.add_argument("-v", "--verbose", action="store_true")
    sub = parser.add_subparsers(dest="command")
    a = sub.add_parser("assign")
    a.add_argument("audio")
    a.add_argument("--transcript", "-t", required=True)
    a.add_argument("--use-embeddings", "-e", action="store_true")
    a.add_argument("--min-trust", default="low", choices=["high", "medium", "low"])
    a.add_argument("--use-llm", "-l", action="store_true")
    a.add_argument("--context", "-c")
    a.add_argument("--expected-speakers")
    a.add_argument("--tags")
    a.add_argument("--threshold", type=float, default=0.3)
    a.add_argument("--output", "-o")
    a.add_argument("--format", "-f", choices=["text", "json"], default="text")
    a.add_argument("--dry-run", "-n", action="store_true")
    a.set_defaults(func=cmd_assign)
    b = sub.add_parser("assign-batch", help="assign many recordings with one resident backend call")
    b.add_argument("manifest", help='JSON list / JSON-lines of {"audio": ..., "transcript": ...}')
    b.add_argument("--min-trust", default="low", choices=["high", "medium", "low"])
    b.add_argument("--tags")
    b.add_argument("--threshold", type=float, default=0.3)
    b.add_argument("--format", "-f", choices=["text", "json"], default="text")
    b.add_argument("--dry-run", "-n", action="store_true")
    b.add_argument("--gpus", type=int, default=1, help="one worker process per GPU; recordings are split over them (no collective)")
    b.add_argument("--shard-bank", action="store_true",
                   help="with --gpus N: shard the bank rows over the GPUs instead (million-profile banks; NCCL all-gather + merge)")
    b.set_defaults(func=cmd_assign_batch)
    sh = sub.add_parser("show", help="Show current assignments for a recording")            # speaker-assign:763-766
    sh.add_argument("audio", help="Path to audio file or b3sum prefix")
    sh.add_argument("--format", "-f", choices=["text", "json", "yaml"], default="text")
    sh.set_defaults(func=cmd_show)
    cl = sub.add_parser("clear", help="Clear assignments for a recording")                  # speaker-assign:769-772
    cl.add_argument("audio", help="Path to audio file or b3sum prefix")
    cl.add_argument("--force", "-f", action="store_true", help="Skip confirmation")
    cl.set_defaults(func=cmd_clear)
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
