#!/usr/bin/env python3
"""Generates tests/golden/*.json by RUNNING THE REFERENCE ITSELF (read-only import from /root/reference).

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Pins:
  * combine_signals_golden.json    -- speaker-assign:418-492 on the 7 survey KATs + 400 seeded random lists
  * embedding_signals_golden.json  -- speaker-assign:262-328 (min-trust filter, default score) with
                                      subprocess.run replaced by canned `speaker_detection identify` JSON
  * assign_embedding_only_golden.json -- min-trust filter + combine_signals on embedding-only rows with
                                      fp32 scores (what sdk_assign / orc_assign restate)
  * transcript_golden.json         -- label discovery / segmentation (speaker-assign:178-246,
                                      transcript.py:123-188) on the config-1 synthetic transcript
  * identify_decorate_golden.json  -- cmd_identify's decoration of backend rows (speaker_detection:1082-1123)
  * trust_golden.json              -- compute_trust_level (speaker_detection:359-379), tag filter (:223-246)
"""
import importlib.machinery
import importlib.util
import io
import json
import os
import random
import sys
import tempfile
import types
from contextlib import redirect_stdout, redirect_stderr
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_script(path: Path, name: str):
    loader = importlib.machinery.SourceFileLoader(name, str(path))
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod


def cfg1_transcript():
    """2 labels, 40 segments, Speechmatics v2 shape."""
    results, t = [], 0.0
    for seg in range(40):
        spk = "S1" if seg % 2 == 0 else "S2"
        for w in range(3):
            results.append({"type": "word", "start_time": round(t, 3), "end_time": round(t + 0.3, 3),
                            "alternatives": [{"content": f"w{seg}_{w}", "confidence": 1.0, "language": "en",
                                              "speaker": spk}]})
            t += 0.35
        results.append({"type": "punctuation", "start_time": round(t - 0.05, 3), "end_time": round(t - 0.05, 3),
                        "attaches_to": "previous", "is_eos": True,
                        "alternatives": [{"content": ".", "confidence": 1.0, "language": "en", "speaker": spk}]})
        t += 0.5
    return {"format": "2.9", "metadata": {"type": "transcription"}, "results": results}


def main():
    sys.path.insert(0, str(REF))
    sa = load_script(REF / "speaker-assign", "ref_speaker_assign")
    sd = load_script(REF / "speaker_detection", "ref_speaker_detection")
    from speaker_detection_backends import transcript as ref_tr

    def run_combine(sigs, thr):
        a = sa.combine_signals("S1", [sa.Signal(t, i, s, dict(ev)) for (t, i, s, ev) in sigs], threshold=thr)
        return {"speaker_id": a.speaker_id, "confidence": a.confidence, "score": a.score,
                "signals": a.signals, "candidates": a.candidates}

    E = lambda i, s, tr=None: ("embedding_match", i, s, ({"trust_level": tr} if tr else {}))
    Cx = lambda i: ("context_expected", i, 0.5, {})
    L = lambda i, s: ("llm_name_detection", i, s, {})
    kats = [
        ([E("alice", .9, "high"), E("bob", .7, "high"), Cx("bob")], 0.3),
        ([E("alice", .8, "low"), E("bob", .5, "high")], 0.1),
        ([E("alice", .91, "high"), E("bob", .62, "medium")], 0.3),
        ([Cx("alice"), Cx("bob"), Cx("carol")], 0.3),
        ([E("zed", .8, "high"), E("amy", .8, "high")], 0.3),
        ([E("x", 1.0, "high"), L("x", 1.0), Cx("x")], 0.3),
        ([E("amy", .9, "invalidated"), E("bob", .9)], 0.1),
        ([], 0.3),
        ([("embedding_match", None, 0.9, {"trust_level": "high"})], 0.3),
    ]
    rng = random.Random(20261018)
    ids = ["alice", "bob", "carol", "dave", "erin", "frank"]
    trusts = ["high", "medium", "low", "invalidated", "unknown", None, "bogus"]
    types_ = ["embedding_match", "embedding_match", "embedding_match", "context_expected", "llm_name_detection",
              "cross_backend_agreement", "made_up_type"]
    for _ in range(400):
        n = rng.randint(0, 8)
        sigs = []
        for _ in range(n):
            t = rng.choice(types_)
            i = rng.choice(ids + [None]) if rng.random() < 0.1 else rng.choice(ids)
            s = rng.choice([rng.random(), round(rng.random(), 2), float(__import__("numpy").float32(rng.random()))])
            ev = {}
            if t == "embedding_match":
                tr = rng.choice(trusts)
                if tr is not None:
                    ev["trust_level"] = tr
            sigs.append((t, i, s, ev))
        kats.append((sigs, rng.choice([0.0, 0.1, 0.2, 0.3, 0.354, 0.5, 0.7])))
    golden = [{"signals": [list(s) for s in sigs], "threshold": thr, "expect": run_combine(sigs, thr)}
              for sigs, thr in kats]
    (OUT / "combine_signals_golden.json").write_text(json.dumps(golden, indent=0))

    # ---- collect_embedding_signals with canned identify output ----
    cases = []
    canned_sets = [
        [{"speaker_id": "alice", "score": 0.91, "trust_level": "high", "embedding_id": "emb-a", "backend": "b200"},
         {"speaker_id": "bob", "score": 0.62, "trust_level": "medium", "embedding_id": "emb-b", "backend": "b200"},
         {"speaker_id": "carol", "score": 0.55, "trust_level": "low", "embedding_id": "emb-c", "backend": "b200"},
         {"speaker_id": "dave", "score": 0.5, "trust_level": "invalidated", "embedding_id": "emb-d", "backend": "b200"},
         {"speaker_id": "erin", "score": 0.45, "trust_level": "unknown", "embedding_id": None, "backend": "b200"},
         {"speaker_id": "frank", "trust_level": "high"},
         {"speaker_id": "", "score": 0.99, "trust_level": "high"},
         {"speaker_id": "gina", "score": 0.4}],
        [],
    ]
    real_run = sa.subprocess.run
    for canned in canned_sets:
        for min_trust in ["low", "medium", "high", "none", "invalidated"]:
            def fake_run(cmd, **kw):
                return types.SimpleNamespace(returncode=0, stdout=json.dumps(canned), stderr="")
            sa.subprocess.run = fake_run
            sigs = sa.collect_embedding_signals("S1", [], Path("/tmp/x.wav"), min_trust=min_trust, tags=None)
            cases.append({"canned": canned, "min_trust": min_trust,
                          "expect": [{"type": s.type, "speaker_id": s.speaker_id, "score": s.score,
                                      "evidence": s.evidence} for s in sigs]})
    sa.subprocess.run = real_run
    (OUT / "embedding_signals_golden.json").write_text(json.dumps(cases, indent=0))

    # ---- embedding-only pipeline: min-trust filter + combine_signals, fp32-representable scores ----
    # This is exactly what the device-side `sdk_assign` (and oracle/canonical.c orc_assign) restate.
    import numpy as _np
    rng2 = random.Random(77)
    eo = []
    for _ in range(400):
        n = rng2.randint(0, 10)
        names = rng2.sample(["spk%02d" % i for i in range(40)], n)
        sims = sorted((float(_np.float32(rng2.uniform(-0.2, 1.0))) for _ in range(n)), reverse=True)
        if n >= 2 and rng2.random() < 0.3:
            sims[1] = sims[0]                      # planted tie
        rows = [{"speaker_id": nm, "score": sc, "trust_level": rng2.choice(["high", "medium", "low", "invalidated", "unknown"]),
                 "embedding_id": "e", "backend": "b200"} for nm, sc in zip(names, sims)]
        min_trust = rng2.choice(["low", "medium", "high"])
        thr = rng2.choice([0.0, 0.1, 0.2, 0.3, 0.35, 0.5])
        def fake_run(cmd, **kw):
            return types.SimpleNamespace(returncode=0, stdout=json.dumps(rows), stderr="")
        sa.subprocess.run = fake_run
        sigs = sa.collect_embedding_signals("S1", [], Path("/tmp/x.wav"), min_trust=min_trust, tags=None)
        a = sa.combine_signals("S1", sigs, threshold=thr)
        eo.append({"rows": rows, "min_trust": min_trust, "threshold": thr,
                   "expect": {"speaker_id": a.speaker_id, "confidence": a.confidence, "score": a.score,
                              "candidates": a.candidates}})
    sa.subprocess.run = real_run
    (OUT / "assign_embedding_only_golden.json").write_text(json.dumps(eo, indent=0))

    # ---- transcript: labels + segments on the config-1 transcript ----
    tr = cfg1_transcript()
    tg = {"transcript": tr, "labels_assign": sa.get_speakers_from_transcript(tr),
          "labels_backend": ref_tr.get_speakers_from_transcript(tr) if hasattr(ref_tr, "get_speakers_from_transcript") else None,
          "segments_assign": {l: sa.get_speaker_segments(tr, l) for l in ["S1", "S2"]},
          "tuples_backend": {l: ref_tr.extract_segments_as_tuples(tr, l) for l in ["S1", "S2"]}}
    aai = {"utterances": [{"speaker": "B", "start": 1500, "end": 2500, "text": "hi"},
                          {"speaker": "A", "start": 0, "end": 1200, "text": "yo"},
                          {"speaker": "B", "start": 2600, "end": 4000, "text": "ok"}]}
    tg["assemblyai"] = {"transcript": aai, "labels_assign": sa.get_speakers_from_transcript(aai),
                        "segments_assign": {l: sa.get_speaker_segments(aai, l) for l in ["A", "B"]},
                        "tuples_backend": {l: ref_tr.extract_segments_as_tuples(aai, l) for l in ["A", "B"]}}
    (OUT / "transcript_golden.json").write_text(json.dumps(tg, indent=0))

    # ---- cmd_identify decoration with a stub backend ----
    import speaker_detection_backends as sdb
    from speaker_detection_backends.base import EmbeddingBackend

    dec_cases = []
    with tempfile.TemporaryDirectory() as td:
        os.environ["SPEAKERS_EMBEDDINGS_DIR"] = td
        db = Path(td) / "db"
        db.mkdir()
        profiles = {
            "alice": {"id": "alice", "version": 1, "names": {"default": "Alice A"}, "tags": ["team"],
                      "embeddings": {"stub": [{"id": "emb-a1", "external_id": None, "trust_level": "high", "created_at": "x"},
                                              {"id": "emb-a2", "external_id": None, "trust_level": "low", "created_at": "x"}]}},
            "bob": {"id": "bob", "version": 1, "names": {"default": "Bob B"}, "tags": [],
                    "embeddings": {"stub": [{"id": "emb-b1", "external_id": None, "trust_level": "medium", "created_at": "x"},
                                            {"id": "emb-b2", "external_id": None, "created_at": "x"}]}},
            "carol": {"id": "carol", "version": 1, "names": {"default": "Carol"}, "tags": ["team"], "embeddings": {}},
        }
        for pid, p in profiles.items():
            (db / f"{pid}.json").write_text(json.dumps(p))
        audio = Path(td) / "a.wav"
        audio.write_bytes(b"RIFF0000WAVE")
        result_sets = [
            [{"speaker_id": "alice", "similarity": 0.91, "embedding_id": "emb-a2", "label": "S1"},
             {"speaker_id": "bob", "confidence": 0.62, "similarity": 0.5},
             {"speaker_id": "ghost", "similarity": 0.4, "embedding_id": "emb-zz"}],
            [],
        ]
        for rs in result_sets:
            class Stub(EmbeddingBackend):
                name = "stub"
                requires_api_key = False
                def enroll_speaker(self, audio_path, segments=None):
                    return {}
                def identify_speaker(self, audio_path, candidates, threshold=0.354):
                    self.seen = [c["id"] for c in candidates]
                    return [dict(r) for r in rs]
            stub = Stub()
            orig = sdb.get_backend
            sdb.get_backend = lambda name: stub
            args = types.SimpleNamespace(audio=str(audio), backend="stub", tags=None, threshold=0.354, format="json")
            so, se = io.StringIO(), io.StringIO()
            with redirect_stdout(so), redirect_stderr(se):
                rc = sd.cmd_identify(args)
            sdb.get_backend = orig
            dec_cases.append({"profiles": profiles, "backend_rows": rs, "rc": rc, "candidates_seen": stub.seen,
                              "stdout_json": json.loads(so.getvalue() or "null"), "stderr": se.getvalue()})
    (OUT / "identify_decorate_golden.json").write_text(json.dumps(dec_cases, indent=0))

    # ---- trust level + tag filter ----
    tcases = []
    for samples in [{}, {"reviewed": ["a"]}, {"unreviewed": ["a"]}, {"reviewed": ["a"], "unreviewed": ["b"]},
                    {"reviewed": ["a"], "rejected": ["c"]}, {"rejected": ["c"]}, {"reviewed": [], "unreviewed": []}]:
        tcases.append({"samples": samples, "expect": sd.compute_trust_level(samples)})
    spk = [{"id": "a", "tags": ["x", "y"]}, {"id": "b", "tags": ["x"]}, {"id": "c", "tags": []}, {"id": "d"}]
    fcases = []
    for tags, any_tag in [(None, False), (["x"], False), (["x", "y"], False), (["x", "y"], True), (["z"], True), ([], False)]:
        fcases.append({"tags": tags, "any_tag": any_tag,
                       "expect": [s["id"] for s in sd.filter_speakers_by_tags(spk, tags, any_tag)]})
    (OUT / "trust_golden.json").write_text(json.dumps({"speakers": spk, "trust": tcases, "filter": fcases}, indent=0))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
