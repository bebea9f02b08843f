#!/usr/bin/env python3
"""Generates tests/golden/schemas_golden.json by RUNNING THE REFERENCE's validators
(speaker_detection_backends/schemas.py:45-251, read-only import from /root/reference) on seeded mutations of valid records.

    python tests/golden/make_schema_golden.py        # authoring container only

Each case: {"kind": "embedding"|"profile", "record": <JSON>, "warnings": [...], "strict_error": <message or null>}."""
import copy
import json
import random
import sys
from pathlib import Path

sys.path.insert(0, "/root/reference")
from speaker_detection_backends.schemas import ValidationError, validate_embedding, validate_profile  # noqa: E402

OUT = Path(__file__).resolve().parent / "schemas_golden.json"

EMB = {"id": "emb-1a2b3c4d", "external_id": "_cas/ab/abcdef.npy", "source_audio": "/x/a.wav", "source_audio_b3sum": "0" * 32,
       "source_segments": [{"start": 0.0, "end": 11.0}], "model_version": "b200-cosine-v1",
       "samples": {"reviewed": ["a" * 32], "unreviewed": [], "rejected": []}, "trust_level": "high",
       "created_at": "2026-01-02T03:04:05.678901+00:00"}
PROF = {"id": "alice", "version": 1, "names": {"default": "Alice"}, "nicknames": [], "description": "", "metadata": {},
        "tags": ["team"], "embeddings": {"b200": [EMB]}, "created_at": "2026-01-01T00:00:00+00:00",
        "updated_at": "2026-01-01T00:00:00+00:00"}

EMB_MUT = [
    ("id", ""), ("id", 7), ("id", None), ("external_id", None), ("external_id", 5), ("external_id", ["x"]),
    ("model_version", "unknown"), ("model_version", 3), ("model_version", None),
    ("trust_level", "invalidated"), ("trust_level", "unknown"), ("trust_level", None), ("trust_level", "HIGH"),
    ("created_at", "2026-01-02T03:04:05Z"), ("created_at", "yesterday"), ("created_at", 12345), ("created_at", None),
    ("samples", None), ("samples", []), ("samples", "x"), ("samples", {"reviewed": "abc"}), ("samples", {"reviewed": [1, 2]}),
    ("samples", {"unreviewed": ["a"], "rejected": [None]}), ("samples", {"other": 1}),
    ("source_segments", None), ("source_segments", "0:1"), ("source_segments", [[0, 1]]), ("source_segments", [{"start": 0}]),
    ("source_segments", [{"start": 0, "end": 1}, {"end": 2}]), ("source_segments", []),
]
PROF_MUT = [
    ("id", ""), ("id", 3), ("names", ["Alice"]), ("names", {}), ("names", {"work": "A"}), ("names", None),
    ("tags", "team"), ("tags", [1, "a"]), ("tags", []), ("embeddings", []), ("embeddings", {"b200": {}}),
    ("embeddings", {"b200": [{"id": "x"}]}), ("embeddings", {"b200": [5, EMB]}), ("embeddings", {}),
    ("embeddings", {"a": [EMB], "b": "no"}), ("version", "1"), ("version", 2.0), ("version", None),
]


def run(kind, rec):
    fn = validate_embedding if kind == "embedding" else validate_profile
    warnings = fn(copy.deepcopy(rec), strict=False)
    try:
        fn(copy.deepcopy(rec), strict=True)
        err = None
    except ValidationError as exc:
        err = str(exc)
    return {"kind": kind, "record": rec, "warnings": warnings, "strict_error": err}


def main():
    rng = random.Random(8)
    cases = [run("embedding", EMB), run("profile", PROF), run("embedding", "nope"), run("profile", [1]), run("embedding", {}),
             run("profile", {})]
    for key, val in EMB_MUT:
        rec = copy.deepcopy(EMB)
        rec[key] = val
        cases.append(run("embedding", rec))
    for key in list(EMB):
        rec = copy.deepcopy(EMB)
        del rec[key]
        cases.append(run("embedding", rec))
    for key, val in PROF_MUT:
        rec = copy.deepcopy(PROF)
        rec[key] = val
        cases.append(run("profile", rec))
    for key in list(PROF):
        rec = copy.deepcopy(PROF)
        del rec[key]
        cases.append(run("profile", rec))
    # several defects at once (order of the warnings is part of the contract)
    for _ in range(60):
        rec = copy.deepcopy(EMB)
        for key, val in rng.sample(EMB_MUT, rng.randint(2, 5)):
            rec[key] = val
        if rng.random() < 0.3:
            del rec[rng.choice(list(rec))]
        cases.append(run("embedding", rec))
    for _ in range(40):
        rec = copy.deepcopy(PROF)
        for key, val in rng.sample(PROF_MUT, rng.randint(2, 4)):
            rec[key] = val
        cases.append(run("profile", rec))
    OUT.write_text(json.dumps(cases, indent=0))
    print(f"{len(cases)} cases -> {OUT}")


if __name__ == "__main__":
    main()
