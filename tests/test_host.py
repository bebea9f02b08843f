"""CPU tests of the host-side mirror of the reference interface (no GPU, no compute calls)."""
import argparse
import ctypes
import io
import json
import os
import sys
import re
from contextlib import redirect_stderr, redirect_stdout
from pathlib import Path

import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native, identify_cli, plugin_api, schemas, signals, store, transcript

ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).parent / "golden"


def load(name):
    return json.loads((GOLD / name).read_text())


# ---- C-ABI: library loads and exports every symbol the header declares ----
def test_abi_exports_every_declared_symbol():
    header = (ROOT / "include" / "sdk_b200.h").read_text()
    declared = set(re.findall(r"\b(sdk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    lib = _native.load()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} not exported"
    assert lib.sdk_abi_version() == 3


def test_no_cpu_fallback():
    """Without a B200 the product path must fail loudly (no oracle, no NumPy behind the API)."""
    if _native.device_count() > 0:
        pytest.skip("a B200 is present")
    with pytest.raises(_native.NativeError) as ei:
        _native.Context(0)
    assert ei.value.code == -19 and "no CPU fallback" in str(ei.value)
    import speaker_diarization_toolkit_b200 as pkg
    for py in Path(pkg.__file__).parent.glob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{py.name} must not touch oracle/"


# ---- signal fusion mirror == reference golden ----
def test_combine_signals_mirror_matches_reference():
    for case in load("combine_signals_golden.json"):
        sigs = [signals.Signal(t, i, s, dict(ev)) for t, i, s, ev in case["signals"]]
        a = signals.combine_signals("S1", sigs, threshold=case["threshold"])
        e = case["expect"]
        assert (a.speaker_id, a.confidence, a.score, a.candidates, a.signals) == \
               (e["speaker_id"], e["confidence"], e["score"], e["candidates"], e["signals"])


def test_signals_from_matches_matches_reference():
    for case in load("embedding_signals_golden.json"):
        got = signals.signals_from_matches(case["canned"], min_trust=case["min_trust"])
        assert [{"type": s.type, "speaker_id": s.speaker_id, "score": s.score, "evidence": s.evidence} for s in got] == case["expect"]


def test_signals_per_label_filter():
    rows = [{"speaker_id": "a", "score": 0.9, "trust_level": "high", "label": "S1"},
            {"speaker_id": "b", "score": 0.8, "trust_level": "high", "label": "S2"},
            {"speaker_id": "c", "score": 0.7, "trust_level": "high"}]
    assert [s.speaker_id for s in signals.signals_from_matches(rows, label="S1")] == ["a", "c"]
    assert [s.speaker_id for s in signals.signals_from_matches(rows, label="S2")] == ["b", "c"]
    assert [s.speaker_id for s in signals.signals_from_matches(rows)] == ["a", "b", "c"]


# ---- transcript mirror == reference golden ----
def test_transcript_mirror_matches_reference():
    g = load("transcript_golden.json")
    tr = g["transcript"]
    assert transcript.get_speakers_from_transcript(tr) == g["labels_assign"] == ["S1", "S2"]
    for lab in ["S1", "S2"]:
        assert transcript.get_speaker_segments(tr, lab) == g["segments_assign"][lab]
        assert [list(t) for t in transcript.extract_segments_as_tuples(tr, lab)] == g["tuples_backend"][lab]
        assert len(g["tuples_backend"][lab]) == 20
    a = g["assemblyai"]
    assert transcript.get_speakers_from_transcript(a["transcript"]) == a["labels_assign"]
    for lab in ["A", "B"]:
        assert transcript.get_speaker_segments(a["transcript"], lab) == a["segments_assign"][lab]
        assert [list(t) for t in transcript.extract_segments_as_tuples(a["transcript"], lab)] == a["tuples_backend"][lab]


# ---- profile store mirror ----
def test_trust_and_tag_filter_match_reference():
    g = load("trust_golden.json")
    for c in g["trust"]:
        assert store.compute_trust_level(c["samples"]) == c["expect"]
    for c in g["filter"]:
        assert [s["id"] for s in store.filter_speakers_by_tags(g["speakers"], c["tags"], c["any_tag"])] == c["expect"]


def _write_profiles(root: Path, profiles: dict):
    (root / "db").mkdir(parents=True, exist_ok=True)
    for pid, p in profiles.items():
        (root / "db" / f"{pid}.json").write_text(json.dumps(p))


def test_cmd_identify_decoration_matches_reference(tmp_path, monkeypatch):
    """Same stdout JSON, stderr status line and rc as the reference's cmd_identify for a stub backend."""
    for case in load("identify_decorate_golden.json"):
        root = tmp_path / f"store{len(case['backend_rows'])}"
        _write_profiles(root, case["profiles"])
        monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(root))
        audio = root / "a.wav"
        audio.write_bytes(b"RIFF0000WAVE")
        rows = [dict(r) for r in case["backend_rows"]]
        for r in rows:
            r.pop("label", None)          # the reference CLI drops unknown keys; compare like for like

        class Stub:
            def identify_speaker(self, audio_path, candidates, threshold):
                self.seen = [c["id"] for c in candidates]
                return rows
        stub = Stub()
        args = argparse.Namespace(audio=str(audio), backend="stub", tags=None, threshold=0.354, format="json")
        so, se = io.StringIO(), io.StringIO()
        with redirect_stdout(so), redirect_stderr(se):
            rc = identify_cli.cmd_identify(args, backend=stub)
        assert rc == case["rc"]
        assert stub.seen == case["candidates_seen"]
        assert json.loads(so.getvalue()) == case["stdout_json"]
        assert se.getvalue() == case["stderr"]


def test_cmd_identify_error_paths(tmp_path, monkeypatch, capsys):
    """Error strings / return codes of speaker_detection:1033-1074 (reference test_cli.py:594-678)."""
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    ns = lambda **kw: argparse.Namespace(**{"backend": "b200", "tags": None, "threshold": 0.354, "format": "json", **kw})
    assert identify_cli.cmd_identify(ns(audio=str(tmp_path / "missing.wav"))) == 1
    assert "Error: Audio file not found" in capsys.readouterr().err
    audio = tmp_path / "a.wav"
    audio.write_bytes(b"x")
    assert identify_cli.cmd_identify(ns(audio=str(audio))) == 1
    assert "No speakers to match against." in capsys.readouterr().err
    _write_profiles(tmp_path, {"al": {"id": "al", "names": {"default": "Al"}, "embeddings": {}}})
    assert identify_cli.cmd_identify(ns(audio=str(audio))) == 1
    assert "No speakers with b200 embeddings." in capsys.readouterr().err
    _write_profiles(tmp_path, {"al": {"id": "al", "names": {"default": "Al"}, "embeddings": {"nope": [{"id": "e"}]}}})
    assert identify_cli.cmd_identify(ns(audio=str(audio), backend="nope")) == 1
    assert "Error loading backend: Unknown backend: nope" in capsys.readouterr().err


def test_bank_build_and_sidecar(tmp_path, monkeypatch):
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    rng = np.random.default_rng(1)
    vecs = {("amy", "emb-1"): rng.standard_normal(8).astype(np.float32), ("amy", "emb-2"): rng.standard_normal(8).astype(np.float32),
            ("bob", "emb-3"): rng.standard_normal(8).astype(np.float32)}
    for (sid, eid), v in vecs.items():
        d = tmp_path / "embeddings" / sid
        d.mkdir(parents=True, exist_ok=True)
        np.save(d / f"{eid}.npy", v)
    handle = store.store_vector_cas(np.ones(8, np.float32))
    cands = [
        {"id": "amy", "embeddings": {"b200": [{"id": "emb-1", "trust_level": "high"}, {"id": "emb-2", "trust_level": "low"}]}},
        {"id": "ghost", "embeddings": {"b200": [{"id": "emb-missing"}]}},
        {"id": "bob", "embeddings": {"b200": [{"id": "emb-3"}, {"id": "emb-x", "external_id": handle, "trust_level": "medium"}]}},
    ]
    snapshot = json.dumps(cands)
    bank = store.build_bank(cands, "b200")
    assert json.dumps(cands) == snapshot, "candidates must not be mutated"
    assert bank.speaker_ids == ["amy", "bob"] and bank.row_speaker.tolist() == [0, 0, 1, 1]
    assert bank.row_trust.tolist() == [0, 2, 4, 1] and bank.row_emb_id == ["emb-1", "emb-2", "emb-3", "emb-x"]
    assert np.array_equal(bank.rows[3], np.ones(8, np.float32))
    with pytest.raises(ValueError, match="dimension"):
        store.build_bank(cands, "b200", dim=16)
    audio = tmp_path / "rec.wav"
    store.save_segment_embeddings(audio, "b200", rng.standard_normal((5, 8)), ["S2", "S1", "S2", "S1", "S10"], [0, 1, 2, 3, 4], [1, 2, 3, 4, 5])
    se = store.load_segment_embeddings(audio, "b200")
    assert se.labels == ["S1", "S10", "S2"] and se.label_index.tolist() == [0, 0, 1, 2, 2] and se.start.tolist() == [1, 3, 4, 0, 2]


# ---- plugin API mirror ----
def test_plugin_registry_and_abc(tmp_path, monkeypatch):
    plugin_api.reload_backends_config()
    monkeypatch.delenv("SPEAKER_BACKENDS_CONFIG", raising=False)
    assert "b200" in plugin_api.list_backends()
    b = plugin_api.get_backend("b200")
    assert b.name == "b200" and b.requires_api_key is False and b.model_version.startswith("b200-")
    assert b.check_embedding_compatibility({"model_version": "b200-cosine-v1"})["compatible"]
    bad = b.check_embedding_compatibility({"model_version": "speechmatics-v2"})
    assert not bad["compatible"] and "Consider re-enrolling" in bad["warning"]
    assert b.get_audio_profile() == plugin_api.AudioProfile()
    with pytest.raises(ValueError, match="Unknown backend: zzz. Available:"):
        plugin_api.get_backend("zzz")
    cfg = tmp_path / "b.yaml"
    cfg.write_text("backends:\n  mine: speaker_diarization_toolkit_b200.backend\n  other:\n    module: json\n")
    monkeypatch.setenv("SPEAKER_BACKENDS_CONFIG", str(cfg))
    plugin_api.reload_backends_config()
    assert plugin_api.list_backends() == ["mine", "other"]
    assert plugin_api.get_backend("mine").name == "b200"
    plugin_api.reload_backends_config()
    assert plugin_api.format_ffmpeg_args(plugin_api.AudioProfile(sample_rate=8000, channels=2, bit_depth=24)) == \
           ["-ar", "8000", "-ac", "2", "-f", "wav", "-acodec", "pcm_s24le"]
    assert plugin_api.format_ffmpeg_args(plugin_api.AudioProfile(format="flac")) == ["-ar", "16000", "-ac", "1", "-f", "flac"]


def test_packed_bank_cache(tmp_path, monkeypatch):
    """the packed cache returns exactly what the per-file layout holds, survives additions,
    edits and deletions, and never wins over a changed .npy."""
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    rng = np.random.default_rng(4)

    def put(sid, eid, vec):
        d = tmp_path / "embeddings" / sid
        d.mkdir(parents=True, exist_ok=True)
        np.save(d / f"{eid}.npy", vec.astype(np.float32))

    cands = []
    for s in range(6):
        recs = []
        for e in range(1 + s % 3):
            put(f"spk{s}", f"emb-{s}{e}", rng.standard_normal(16))
            recs.append({"id": f"emb-{s}{e}", "trust_level": ["high", "medium", "low"][e]})
        cands.append({"id": f"spk{s}", "embeddings": {"b200": recs}})

    def same(a, b):
        return (np.array_equal(a.rows, b.rows) and np.array_equal(a.row_speaker, b.row_speaker) and np.array_equal(a.row_trust, b.row_trust)
                and a.row_emb_id == b.row_emb_id and a.speaker_ids == b.speaker_ids)

    plain = store.build_bank(cands, "b200")
    assert same(store.build_bank_cached(cands, "b200"), plain)               # cold: builds the pack
    pack = tmp_path / "embeddings" / ".bank-b200-D16.f32"
    assert pack.exists() and pack.stat().st_size == plain.P * 16 * 4
    reads = []
    real_load = np.load
    monkeypatch.setattr(np, "load", lambda p, *a, **k: (reads.append(str(p)), real_load(p, *a, **k))[1])
    assert same(store.build_bank_cached(cands, "b200"), plain)
    assert [r for r in reads if r.endswith(".npy")] == []                    # warm: no vector file opened (only the binary index)
    # a changed vector is re-read (mtime/size), a new record is appended, a removed speaker disappears
    import os as _os, time as _time
    put("spk2", "emb-20", np.arange(16))
    _os.utime(tmp_path / "embeddings" / "spk2" / "emb-20.npy", ns=(_time.time_ns(), _time.time_ns() + 10**9))
    put("spk9", "emb-90", rng.standard_normal(16))
    cands2 = [c for c in cands if c["id"] != "spk0"] + [{"id": "spk9", "embeddings": {"b200": [{"id": "emb-90"}]}}]
    reads.clear()
    got = store.build_bank_cached(cands2, "b200")
    monkeypatch.setattr(np, "load", real_load)
    assert same(got, store.build_bank(cands2, "b200"))
    assert sorted(Path(r).name for r in reads if r.endswith(".npy")) == ["emb-20.npy", "emb-90.npy"]
    assert np.array_equal(got.rows[got.row_emb_id.index("emb-20")], np.arange(16, dtype=np.float32))


def test_packed_bank_cache_trust_mode_and_dimension_change(tmp_path, monkeypatch):
    """SPEAKER_B200_BANK_CACHE=trust takes pack rows by key without a stat() per file; a store re-enrolled at another
    dimension gets its own pack instead of failing against the stale one; the index is a binary .npz."""
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    rng = np.random.default_rng(9)

    def put(sid, eid, vec):
        d = tmp_path / "embeddings" / sid
        d.mkdir(parents=True, exist_ok=True)
        np.save(d / f"{eid}.npy", np.asarray(vec, np.float32))

    cands = []
    for s in range(5):
        put(f"spk{s}", f"e{s}", rng.standard_normal(8))
        cands.append({"id": f"spk{s}", "embeddings": {"b200": [{"id": f"e{s}", "trust_level": "high"}]}})
    plain = store.build_bank(cands, "b200")
    assert np.array_equal(store.build_bank_cached(cands, "b200").rows, plain.rows)
    assert (tmp_path / "embeddings" / ".bank-b200-D8.idx.npz").exists() and not list((tmp_path / "embeddings").glob("*.tmp*"))
    monkeypatch.setenv("SPEAKER_B200_BANK_CACHE", "trust")
    stats = []
    real_stat = Path.stat
    monkeypatch.setattr(Path, "stat", lambda self, **k: (stats.append(self.name), real_stat(self, **k))[1])
    got = store.build_bank_cached(cands, "b200")
    monkeypatch.setattr(Path, "stat", real_stat)
    assert np.array_equal(got.rows, plain.rows) and got.speaker_ids == plain.speaker_ids
    assert not [n for n in stats if n.endswith(".npy")]                       # no per-vector stat in trust mode
    # a new speaker in trust mode: only that vector is read, the pack grows by one row
    put("spk9", "e9", rng.standard_normal(8))
    cands9 = cands + [{"id": "spk9", "embeddings": {"b200": [{"id": "e9"}]}}]
    got = store.build_bank_cached(cands9, "b200")
    assert np.array_equal(got.rows, store.build_bank(cands9, "b200").rows)
    assert (tmp_path / "embeddings" / ".bank-b200-D8.f32").stat().st_size == 6 * 8 * 4
    monkeypatch.delenv("SPEAKER_B200_BANK_CACHE")
    assert np.array_equal(store.build_bank_cached(cands9, "b200").rows, got.rows)
    # re-enrolled at D = 12: the D8 pack must not dictate the dimension
    cands12 = []
    for s in range(3):
        put(f"spk{s}", f"e{s}", rng.standard_normal(12))
        cands12.append({"id": f"spk{s}", "embeddings": {"b200": [{"id": f"e{s}"}]}})
    got12 = store.build_bank_cached(cands12, "b200")
    assert got12.rows.shape == (3, 12) and np.array_equal(got12.rows, store.build_bank(cands12, "b200").rows)
    assert (tmp_path / "embeddings" / ".bank-b200-D12.f32").exists()


def test_profile_pack_equals_the_per_file_listing(tmp_path, monkeypatch):
    """list_all_speakers_cached: same list and order as the reference's per-file pass; only changed files are parsed again"""
    import builtins
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    (tmp_path / "db").mkdir()
    for i in (3, 1, 2):
        (tmp_path / "db" / f"spk{i}.json").write_text(json.dumps({"id": f"spk{i}", "tags": ["t"] if i == 2 else []}))
    (tmp_path / "db" / "broken.json").write_text("{nope")
    plain = store.list_all_speakers()
    assert [p["id"] for p in plain] == ["spk1", "spk2", "spk3"]
    assert store.list_all_speakers_cached() == plain and (tmp_path / "db" / ".profiles.pack").exists()
    assert store.list_all_speakers() == plain                         # the pack is invisible to the per-file glob
    opened = []
    real_open = builtins.open
    monkeypatch.setattr(builtins, "open", lambda f, *a, **k: (opened.append(str(f)), real_open(f, *a, **k))[1])
    assert store.list_all_speakers_cached() == plain
    assert [o for o in opened if o.endswith(".json") and "spk" in o] == []       # warm: no profile file opened
    opened.clear()
    monkeypatch.setattr(builtins, "open", real_open)
    import os as _os, time as _time
    (tmp_path / "db" / "spk2.json").write_text(json.dumps({"id": "spk2", "tags": ["changed"]}))
    _os.utime(tmp_path / "db" / "spk2.json", ns=(_time.time_ns(), _time.time_ns() + 10**9))
    (tmp_path / "db" / "spk0.json").write_text(json.dumps({"id": "spk0"}))
    (tmp_path / "db" / "spk3.json").unlink()
    got = store.list_all_speakers_cached()
    assert got == store.list_all_speakers() and [p["id"] for p in got] == ["spk0", "spk1", "spk2"] and got[2]["tags"] == ["changed"]
    monkeypatch.setenv("SPEAKER_B200_PROFILE_CACHE", "0")
    assert store.list_all_speakers_cached() == got


# ---- enrollment / bank-writing path ------------------------------------------------------------
ENROLL_RECORD_KEYS = ["id", "external_id", "source_audio", "source_audio_b3sum", "source_segments", "model_version", "samples",
                      "trust_level", "created_at"]        # speaker_detection:890-901, in this order


def _enroll_store(tmp_path, D=24, n=30, seed=5):
    rng = np.random.default_rng(seed)
    (tmp_path / "db").mkdir(parents=True)
    prof = {"id": "alice", "version": 1, "names": {"default": "Alice"}, "nicknames": [], "description": "", "metadata": {},
            "tags": [], "embeddings": {}, "created_at": "2026-01-01T00:00:00+00:00", "updated_at": "2026-01-01T00:00:00+00:00"}
    (tmp_path / "db" / "alice.json").write_text(json.dumps(prof))
    audio = tmp_path / "a.wav"
    audio.write_bytes(b"RIFF" + bytes(40) + b"enroll")
    emb = rng.standard_normal((n, D)).astype(np.float32) * rng.uniform(0.5, 9, size=(n, 1)).astype(np.float32)
    labels = ["S1" if i % 3 else "S2" for i in range(n)]
    start = np.arange(n) * 2.0
    store.save_segment_embeddings(audio, "b200", emb, labels, start, start + 1.5)
    return audio, emb, labels, start


def test_enroll_cli_writes_reference_record_and_vector(tmp_path, monkeypatch, capsys):
    from speaker_diarization_toolkit_b200 import identify_cli
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    monkeypatch.delenv("SPEAKER_BACKENDS_CONFIG", raising=False)
    audio, emb, labels, start = _enroll_store(tmp_path)
    # errors first, with the reference's messages and rc (speaker_detection:759-780)
    assert identify_cli.main(["enroll", "bob", str(audio)]) == 1
    assert "Error: Speaker 'bob' not found. Use 'add' first." in capsys.readouterr().err
    assert identify_cli.main(["enroll", "alice", str(tmp_path / "nope.wav")]) == 1
    assert "Error: Audio file not found:" in capsys.readouterr().err
    assert identify_cli.main(["enroll", "alice", str(audio), "-s", "5:3"]) == 1
    assert "Error: Invalid segment '5:3'. Start must be < end." in capsys.readouterr().err
    assert identify_cli.main(["enroll", "alice", str(audio), "-s", "0:11", "--dry-run"]) == 0
    out = capsys.readouterr().out
    assert "Would enroll speaker: alice" in out and "Segments: 1 (11.0s total)" in out
    # the real thing: segments overlapping [0, 11) s = sidecar rows 0..5
    assert identify_cli.main(["-q", "enroll", "alice", str(audio), "-b", "b200", "-s", "0:11", "--trust-level", "high"]) == 0
    cap = capsys.readouterr()
    assert "Enrolled embedding emb-" in cap.out and "(no samples tracked)" in cap.out and cap.err == ""
    prof = store.load_speaker("alice")
    (rec,) = prof["embeddings"]["b200"]
    assert list(rec.keys()) == ENROLL_RECORD_KEYS
    assert rec["trust_level"] == "high" and rec["model_version"] == "b200-cosine-v1" and rec["source_segments"] == [{"start": 0.0, "end": 11.0}]
    assert rec["samples"] == {"reviewed": [], "unreviewed": [], "rejected": []}
    x = emb[:6].astype(np.float64)
    want = (x / np.linalg.norm(x, axis=1, keepdims=True)).mean(axis=0).astype(np.float32)
    canonical = tmp_path / "embeddings" / "alice" / f"{rec['id']}.npy"
    assert np.array_equal(np.load(canonical), want)
    assert np.array_equal(np.load(tmp_path / "embeddings" / rec["external_id"]), want)      # content-addressed handle
    bank = store.build_bank_cached([prof], "b200")
    assert bank.P == 1 and np.array_equal(bank.rows[0], want) and bank.row_emb_id == [rec["id"]]
    # --from-transcript picks the label's segments; default trust (no samples) is "low" (speaker_detection:378-379)
    words = []
    for i, lab in enumerate(labels):
        words.append({"type": "word", "start_time": float(start[i]), "end_time": float(start[i]) + 1.5,
                      "alternatives": [{"content": "w", "confidence": 1.0, "speaker": lab}]})
    tpath = tmp_path / "a.json"
    tpath.write_text(json.dumps({"format": "2.9", "results": words}))
    assert identify_cli.main(["enroll", "alice", str(audio), "-t", str(tpath)]) == 1
    assert "--speaker-label required with --from-transcript" in capsys.readouterr().err
    assert identify_cli.main(["enroll", "alice", str(audio), "-t", str(tpath), "-l", "S2"]) == 0
    capsys.readouterr()
    recs = store.load_speaker("alice")["embeddings"]["b200"]
    assert len(recs) == 2 and recs[1]["trust_level"] == "low"
    sel = np.asarray([l == "S2" for l in labels])
    x = emb[sel].astype(np.float64)
    want2 = (x / np.linalg.norm(x, axis=1, keepdims=True)).mean(axis=0).astype(np.float32)
    assert np.array_equal(np.load(tmp_path / "embeddings" / "alice" / f"{recs[1]['id']}.npy"), want2)
    # a recording whose embeddings have another dimension is rejected before anything is stored
    audio2 = tmp_path / "b.wav"
    audio2.write_bytes(b"RIFF" + bytes(40) + b"other")
    store.save_segment_embeddings(audio2, "b200", np.ones((3, 16), np.float32), ["S1"] * 3)
    assert identify_cli.main(["enroll", "alice", str(audio2)]) == 1
    assert "Error during enrollment: embedding is 16-d but the bank enrolled for b200 is 24-d" in capsys.readouterr().err
    assert len(store.load_speaker("alice")["embeddings"]["b200"]) == 2


@pytest.mark.skipif(not Path("/root/reference/speaker_detection").exists(), reason="needs the reference tree (authoring container)")
def test_enroll_through_the_reference_cli_with_the_b200_backend(tmp_path):
    """The UNMODIFIED reference `speaker_detection enroll` with this package plugged in as a backend produces a record our
    store resolves to the same vector (via the content-addressed external_id), and our `enroll` writes the same keys."""
    import subprocess
    audio, emb, labels, start = _enroll_store(tmp_path)
    cfg = tmp_path / "backends.yaml"
    cfg.write_text("backends:\n  b200:\n    module: speaker_diarization_toolkit_b200.backend\n")
    env = dict(os.environ, SPEAKERS_EMBEDDINGS_DIR=str(tmp_path), SPEAKER_BACKENDS_CONFIG=str(cfg), PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "/root/reference/speaker_detection", "enroll", "alice", str(audio), "-b", "b200", "-s", "0:11"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    os.environ["SPEAKERS_EMBEDDINGS_DIR"] = str(tmp_path)
    try:
        prof = store.load_speaker("alice")
        (rec,) = prof["embeddings"]["b200"]
        assert list(rec.keys()) == ENROLL_RECORD_KEYS
        x = emb[:6].astype(np.float64)
        want = (x / np.linalg.norm(x, axis=1, keepdims=True)).mean(axis=0).astype(np.float32)
        bank = store.build_bank([prof], "b200")
        assert bank.P == 1 and np.array_equal(bank.rows[0], want)
    finally:
        os.environ.pop("SPEAKERS_EMBEDDINGS_DIR", None)


# ---- schemas: the validators the new records must pass -------------------------------------------
def test_schema_validators_match_the_reference_golden():
    cases = load("schemas_golden.json")
    assert len(cases) >= 170
    n_err = 0
    for c in cases:
        fn = schemas.validate_embedding if c["kind"] == "embedding" else schemas.validate_profile
        assert fn(c["record"]) == c["warnings"], c["record"]
        if c["strict_error"] is None:
            assert fn(c["record"], strict=True) == c["warnings"]
        else:
            n_err += 1
            with pytest.raises(schemas.ValidationError) as exc:
                fn(c["record"], strict=True)
            assert str(exc.value) == c["strict_error"]
    assert n_err > 50


def test_enrolled_records_pass_the_validators(tmp_path, monkeypatch, capsys):
    """What `enroll` writes is valid by the mirror and, where the reference tree is present, by the reference itself."""
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    monkeypatch.delenv("SPEAKER_BACKENDS_CONFIG", raising=False)
    audio, *_ = _enroll_store(tmp_path)
    assert identify_cli.main(["-q", "enroll", "alice", str(audio), "-s", "0:11"]) == 0
    capsys.readouterr()
    prof = store.load_speaker("alice")
    assert schemas.validate_profile(prof, strict=True) == []
    assert schemas.validate_embedding(prof["embeddings"]["b200"][0], strict=True) == []
    if Path("/root/reference/speaker_detection_backends/schemas.py").exists():
        sys.path.insert(0, "/root/reference")
        try:
            from speaker_detection_backends import schemas as ref_schemas
            assert ref_schemas.validate_profile(prof, strict=True) == []
        finally:
            sys.path.remove("/root/reference")


def test_validate_cli_matches_the_reference(tmp_path, monkeypatch, capsys):
    """`speaker_detection validate` (speaker_detection:1307-1361): same stdout and return code as the reference CLI on a
    store with a clean, a flawed and a badly flawed profile (reference run only where its tree is present)."""
    import subprocess
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    (tmp_path / "db").mkdir()
    good = {"id": "alice", "version": 1, "names": {"default": "Alice"}, "tags": ["a"], "embeddings": {"b200": [
        {"id": "emb-1", "external_id": None, "created_at": "2026-01-01T00:00:00+00:00", "model_version": "b200-cosine-v1", "trust_level": "high"}]}}
    # (every profile carries the current schema version: the reference migrates older ones in place when it loads them)
    flawed = {"id": "bob", "version": 1, "names": {"work": "Bob"}, "tags": ["x"], "embeddings": {"b200": [
        {"id": "emb-2", "external_id": 7, "created_at": "yesterday", "model_version": "unknown", "trust_level": "unknown"}]}}
    bad = {"id": "carol", "version": 1, "names": ["Carol"], "tags": "t", "embeddings": {"b200": {}}}
    for prof in (good, flawed, bad):
        (tmp_path / "db" / f"{prof['id']}.json").write_text(json.dumps(prof))
    runs = [["validate"], ["validate", "-v"], ["validate", "--strict"], ["validate", "bob", "--strict"], ["validate", "alice", "--strict", "-v"],
            ["validate", "-q", "--strict"], ["validate", "nobody"]]
    for argv in runs:
        rc = identify_cli.main(argv)
        cap = capsys.readouterr()
        if argv == ["validate"]:
            assert rc == 0 and "Validated 3 profiles" in cap.out and "2 profiles with issues" in cap.out
            assert "  - embeddings.b200[0]: Embedding 'external_id' must be a string or null, got int" in cap.out
        if argv == ["validate", "nobody"]:
            assert rc == 1 and "Error: Speaker 'nobody' not found." in cap.err
        if Path("/root/reference/speaker_detection").exists():
            r = subprocess.run([sys.executable, "/root/reference/speaker_detection", *argv], capture_output=True, text=True,
                               env=dict(os.environ, SPEAKERS_EMBEDDINGS_DIR=str(tmp_path)))
            assert (r.returncode, r.stdout, r.stderr) == (rc, cap.out, cap.err), argv


def test_assign_show_and_clear_match_the_reference(tmp_path, monkeypatch, capsys):
    """`speaker-assign show|clear` (speaker-assign:652-728): same stdout / stderr / return codes as the reference CLI on an
    assignments file of the shape `assign` writes."""
    import subprocess
    from speaker_diarization_toolkit_b200 import assign_cli
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    audio = tmp_path / "a.wav"
    audio.write_bytes(b"RIFF" + bytes(40) + b"show")
    b3 = assign_cli.compute_b3sum(audio)
    data = {"schema_version": 1, "recording_b3sum": b3, "transcript_path": "/x/t.json", "assigned_at": "2026-01-01T00:00:00Z",
            "method": "speaker-assign-v1.0.0", "context": None, "min_trust": "low", "threshold": 0.3,
            "mappings": {"S1": {"speaker_id": "alice", "confidence": "low", "score": 0.364,
                                "signals": [{"type": "embedding_match", "score": 0.91, "embedding_id": "emb-1", "trust_level": "high", "backend": "b200"}],
                                "candidates": [{"speaker_id": "bob", "score": 0.1736}]},
                         "S2": {"speaker_id": None, "confidence": "unassigned", "score": 0.0, "signals": []}}}
    adir = tmp_path / "assignments"
    adir.mkdir()
    ref_ok = Path("/root/reference/speaker-assign").exists()

    def both(argv, restore=False):
        if restore:
            assign_cli._save_yaml(adir / f"{b3}.yaml", data)
        rc = assign_cli.main(argv)
        cap = capsys.readouterr()
        if ref_ok:
            if restore:
                assign_cli._save_yaml(adir / f"{b3}.yaml", data)
            r = subprocess.run([sys.executable, "/root/reference/speaker-assign", *argv], capture_output=True, text=True,
                               env=dict(os.environ, SPEAKERS_EMBEDDINGS_DIR=str(tmp_path)))
            assert (r.returncode, r.stdout, r.stderr) == (rc, cap.out, cap.err), argv
        return rc, cap

    rc, cap = both(["show", str(audio)])
    assert rc == 1 and "Error: No assignments found for this recording" in cap.err
    assign_cli._save_yaml(adir / f"{b3}.yaml", data)
    rc, cap = both(["show", str(audio)])
    assert rc == 0 and "  S1 -> alice" in cap.out and "  S2 -> (unassigned)" in cap.out and "candidates: bob(0.17)" in cap.out
    for fmt in ("json", "yaml"):
        rc, cap = both(["show", str(audio), "--format", fmt])
        assert rc == 0
    assert json.loads(both(["show", str(audio), "-f", "json"])[1].out) == data
    rc, cap = both(["show", str(tmp_path / "nope.wav")])
    assert rc == 1 and "Error: Could not resolve audio:" in cap.err
    rc, cap = both(["clear", str(audio), "--force"], restore=True)
    assert rc == 0 and f"Cleared assignments: {b3[:8]}..." in cap.out and not (adir / f"{b3}.yaml").exists()
    rc, cap = both(["clear", str(audio), "--force"])
    assert rc == 0 and "No assignments found for this recording" in cap.err


def test_phase_trace_build_of_the_small_query_kernels_compiles(tmp_path):
    """tools/gv_trace.py: the diagnostic build (-DSDK_GV_TRACE: clock stamps at the phase boundaries of k_gemv8 /
    k_gemv8_tail) must keep compiling next to the shipped one; only the object is built here, into a scratch directory."""
    import shutil
    import subprocess
    from speaker_diarization_toolkit_b200 import build as b
    if not (shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists()):
        pytest.skip("no nvcc")
    obj = tmp_path / "gemv_trace.o"
    r = subprocess.run([b.nvcc(), *b.NVCC_FLAGS, "-DSDK_GV_TRACE", "-c", str(b.CSRC / "gemv.cu"), "-o", str(obj)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    sym = subprocess.run(["nm", str(obj)], capture_output=True, text=True).stdout
    assert "sdk_debug_gv_trace" in sym
    shipped = subprocess.run(["nm", "-D", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sdk_debug_gv_trace" not in shipped          # the shipped library holds none of it
