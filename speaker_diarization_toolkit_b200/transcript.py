"""What a "label" and a "segment" are (boundary helper).

Mirrors the two segmenters of the reference -- pinned by tests/golden/transcript_golden.json:
  * speaker-assign:169-246      `get_speakers_from_transcript`, `get_speaker_segments`
  * speaker_detection_backends/transcript.py:25-53, :123-188   `detect_transcript_format`,
                                `extract_segments_as_tuples` (the one the backend ABC exposes, base.py:182-200;
                                canonical segment order of the embedding sidecar)
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Tuple


def load_transcript(path) -> Dict[str, Any]:
    with open(path, "r") as fh:
        return json.load(fh)


def detect_transcript_format(data: Dict[str, Any]) -> str:
    """transcript.py:25-53: AssemblyAI has `utterances`; Speechmatics a `results` list of word items."""
    if "utterances" in data:
        return "assemblyai"
    results = data.get("results")
    if isinstance(results, list) and results:
        head = results[0]
        if "alternatives" in head or "start_time" in head or head.get("type") in ("word", "punctuation"):
            return "speechmatics"
    return "unknown"


def _assign_format(data: Dict[str, Any]) -> str:
    """speaker-assign:169-175 sniffs more loosely than transcript.py."""
    if "utterances" in data:
        return "assemblyai"
    return "speechmatics" if "results" in data else "unknown"


def _item_speaker_assign(item: Dict[str, Any]):
    """speaker-assign:215-226: the last alternative carrying a speaker wins, a top-level speaker overrides."""
    spk, text = None, ""
    for alt in item.get("alternatives", []):
        if alt.get("speaker"):
            spk = alt["speaker"]
        if alt.get("content"):
            text = alt["content"]
    if item.get("speaker"):
        spk = item["speaker"]
    return spk, text


def get_speakers_from_transcript(data: Dict[str, Any]) -> List[str]:
    """Sorted unique labels (speaker-assign:178-196)."""
    found = set()
    fmt = _assign_format(data)
    if fmt == "assemblyai":
        found.update(u["speaker"] for u in data.get("utterances", []) if u.get("speaker"))
    elif fmt == "speechmatics":
        for item in data.get("results", []):
            found.update(a["speaker"] for a in item.get("alternatives", []) if a.get("speaker"))
            if item.get("speaker"):
                found.add(item["speaker"])
    return sorted(found)


def get_speaker_segments(data: Dict[str, Any], speaker_label: str) -> List[Dict[str, Any]]:
    """Per-label segment dicts {start,end,text} (speaker-assign:199-246): consecutive same-speaker results merge;
    any other result (including one without a start_time) closes the open segment."""
    fmt = _assign_format(data)
    out: List[Dict[str, Any]] = []
    if fmt == "assemblyai":
        for u in data.get("utterances", []):
            if u.get("speaker") == speaker_label:
                out.append({"start": u.get("start", 0) / 1000.0, "end": u.get("end", 0) / 1000.0, "text": u.get("text", "")})
        return out
    if fmt != "speechmatics":
        return out
    cur = None
    for item in data.get("results", []):
        spk, text = _item_speaker_assign(item)
        if spk == speaker_label and item.get("start_time") is not None:
            end = item.get("end_time", item["start_time"])
            if cur is None:
                cur = {"start": item["start_time"], "end": end, "text": text}
            else:
                cur["end"] = end
                if text:
                    cur["text"] += " " + text
        elif cur is not None:
            out.append(cur)
            cur = None
    if cur is not None:
        out.append(cur)
    return out


def extract_segments_as_tuples(data: Dict[str, Any], speaker_label: str) -> List[Tuple[float, float]]:
    """(start,end) tuples without merging (transcript.py:123-188): only `word` items count; an item without
    a speaker is 'UU'; a run of the label's words is one segment."""
    fmt = detect_transcript_format(data)
    spans: List[Tuple[float, float]] = []
    if fmt == "assemblyai":
        return [(u.get("start", 0) / 1000.0, u.get("end", 0) / 1000.0)
                for u in data.get("utterances", []) if u.get("speaker") == speaker_label]
    if fmt != "speechmatics":
        return spans
    run_start = run_end = None
    prev = None
    for item in data.get("results", []):
        if item.get("type") != "word":
            continue
        spk = item.get("speaker")
        if not spk:
            alts = item.get("alternatives", [])
            spk = alts[0].get("speaker") if alts else None
        spk = spk or "UU"
        if spk == speaker_label:
            if prev != speaker_label:
                run_start = item.get("start_time", 0)
            run_end = item.get("end_time", 0)
        elif prev == speaker_label and run_start is not None:
            spans.append((run_start, run_end))
            run_start = None
        prev = spk
    if prev == speaker_label and run_start is not None:
        spans.append((run_start, run_end))
    return spans
