"""`speaker-assign assign` without the embedding step (speaker-assign:497-650): the host side of row a8 -- transcript
labels, catalog / --expected-speakers context signals, combine_signals, the YAML record, every output mode and error --
run side by side with the REFERENCE CLI on the same store (when /root/reference is there: this container; the GPU box
runs only the `-m gpu` tests) and compared byte for byte, the `assigned_at` timestamp aside.  The cases are the ones of
the reference's own evals/speaker_detection/test_speaker_assign.py (:206-1049).  No device is touched: without `-e`
the command never opens a context."""
import json
import os
import re
import struct
import subprocess
import sys
import hashlib
from pathlib import Path

import pytest
import yaml

from speaker_diarization_toolkit_b200 import assign_cli

REF = Path("/root/reference/speaker-assign")
STAMP = re.compile(r"\d{4}-\d{2}-\d{2}T\d{2}:\d{2}:\d{2}(\.\d+)?(Z|\+00:00)")


def wav(path: Path, unique: str, seconds: float = 0.05) -> Path:
    """a small PCM file whose content (hence its b3sum) depends on `unique` (test_speaker_assign.py:50-103)"""
    n = int(44100 * seconds) * 2
    seed = hashlib.sha256(unique.encode()).digest()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + n) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 44100, 88200, 2, 16))
        f.write(b"data" + struct.pack("<I", n) + (seed * (n // len(seed) + 1))[:n])
    return path


def transcript(path: Path, speakers: int = 2) -> Path:
    """AssemblyAI-style utterances (test_speaker_assign.py:106-143)"""
    utt = {2: [("A", "Hello everyone, this is Alice speaking"), ("B", "Hi Alice, Bob here"), ("A", "How is the project going?"),
               ("B", "Making good progress, thanks for asking"), ("A", "Great, let me know if you need any help")],
           3: [("A", "Hello everyone, I'm Alice"), ("B", "Hi, Bob here"), ("C", "And I'm Carol"), ("A", "Let's start the meeting"),
               ("B", "Sounds good"), ("C", "I have some updates")],
           1: [("A", "Hello there")]}[speakers]
    data = {"utterances": [{"speaker": s, "start": 1000 + 5000 * i, "end": 5000 + 5000 * i, "text": t} for i, (s, t) in enumerate(utt)]}
    path.write_text(json.dumps(data, indent=2))
    return path


def transcript_speechmatics(path: Path) -> Path:
    """test_speaker_assign.py:146-159"""
    rows = [(1.0, 2.0, "S1", "Hello"), (2.5, 3.5, "S2", "Hi there"), (4.0, 5.0, "S1", "How are you")]
    path.write_text(json.dumps({"results": [{"start_time": a, "end_time": b, "speaker": s, "alternatives": [{"content": c, "speaker": s}]}
                                            for a, b, s, c in rows]}, indent=2))
    return path


def catalog_entry(root: Path, audio: Path, name, expected) -> str:
    """test_speaker_assign.py:162-201"""
    b3 = assign_cli.compute_b3sum(audio)
    (root / "catalog").mkdir(exist_ok=True)
    (root / "catalog" / f"{b3}.yaml").write_text(yaml.dump({"recording": {"b3sum": b3, "original_path": str(audio)},
                                                            "context": {"name": name, "expected_speakers": expected or []}},
                                                           default_flow_style=False))
    return b3


def scrub(text: str) -> str:
    return STAMP.sub("<T>", text)


class Pair:
    """runs one argv through this repo's CLI (in process) and through the reference's (subprocess), each on its own copy
    of the store so that what one saves cannot be seen by the other; compares return code, stdout, stderr and the files"""

    def __init__(self, tmp_path, monkeypatch, capsys):
        self.ours, self.theirs = tmp_path / "ours", tmp_path / "theirs"
        self.ours.mkdir()
        self.theirs.mkdir()
        self.monkeypatch, self.capsys = monkeypatch, capsys
        self.ref_ok = REF.exists()
        self.extra = {}                 # environment both sides get on top of SPEAKERS_EMBEDDINGS_DIR

    def both(self, build, argv_of):
        """build(root) -> dict of paths; argv_of(paths) -> argv.  Returns (rc, out, err, ours_root, paths)."""
        paths = build(self.ours)
        self.monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(self.ours))
        for key, val in self.extra.items():
            self.monkeypatch.setenv(key, val)
        rc = assign_cli.main(argv_of(paths))
        cap = self.capsys.readouterr()
        if self.ref_ok:
            rpaths = build(self.theirs)
            r = subprocess.run([sys.executable, str(REF), *argv_of(rpaths)], capture_output=True, text=True,
                               env=dict(os.environ, **self.extra, SPEAKERS_EMBEDDINGS_DIR=str(self.theirs)))
            swap = lambda s: scrub(s).replace(str(self.theirs), "<ROOT>")
            mine = lambda s: scrub(s).replace(str(self.ours), "<ROOT>")
            assert (r.returncode, swap(r.stdout), swap(r.stderr)) == (rc, mine(cap.out), mine(cap.err)), argv_of(paths)
            ours_files = sorted(p.relative_to(self.ours) for p in self.ours.rglob("*.yaml"))
            their_files = sorted(p.relative_to(self.theirs) for p in self.theirs.rglob("*.yaml"))
            assert ours_files == their_files
            for rel in ours_files:
                assert mine((self.ours / rel).read_text()) == swap((self.theirs / rel).read_text()), rel
        return rc, cap.out, cap.err, paths


@pytest.fixture()
def pair(tmp_path, monkeypatch, capsys):
    return Pair(tmp_path, monkeypatch, capsys)


def two(root, speakers=2, uid="basic"):
    return {"audio": wav(root / "meeting.wav", uid), "t": transcript(root / "transcript.json", speakers), "root": root}


def saved(paths):
    b3 = assign_cli.compute_b3sum(paths["audio"])
    return yaml.safe_load((paths["root"] / "assignments" / f"{b3}.yaml").read_text())


def test_assign_basic_and_three_speakers(pair):
    """:206, :242 -- no signals at all: every label is there, unassigned, and the record is saved"""
    rc, out, err, paths = pair.both(two, lambda p: ["assign", str(p["audio"]), "-t", str(p["t"])])
    assert rc == 0 and "Found 2 speakers: A, B" in out and "Assigned: 0/2" in out and "  A -> (unassigned) (unassigned, score: 0.00)" in out
    rec = saved(paths)
    assert rec["schema_version"] == 1 and set(rec["mappings"]) == {"A", "B"} and rec["method"].startswith("speaker-assign-v")
    assert rec["mappings"]["A"] == {"speaker_id": None, "confidence": "unassigned", "score": 0.0, "signals": []}
    rc, out, err, paths = pair.both(lambda r: two(r, 3, "three"), lambda p: ["assign", str(p["audio"]), "--transcript", str(p["t"])])
    assert rc == 0 and "Found 3 speakers: A, B, C" in out and set(saved(paths)["mappings"]) == {"A", "B", "C"}


def test_assign_expected_speakers_from_the_catalog_and_from_the_command_line(pair):
    """:280, :320 -- context signals (speaker-assign:331-357): every expected speaker is a weak candidate for every label"""
    def with_catalog(root):
        p = two(root, 2, "catalog")
        catalog_entry(root, p["audio"], "team-standup", ["alice", "bob"])
        return p
    rc, out, err, paths = pair.both(with_catalog, lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "-f", "json"])
    rec = json.loads(out[out.index("{"):])
    assert rc == 0 and rec["context"] == "team-standup"
    for label in ("A", "B"):
        m = rec["mappings"][label]        # 0.2 x 0.5 per expected speaker: below the threshold, both are candidates, alice first
        assert m["speaker_id"] is None and m["score"] == 0.1 and [c["speaker_id"] for c in m["candidates"]] == ["alice", "bob"]
        assert m["signals"] == [{"type": "context_expected", "score": 0.5, "context": "team-standup", "reason": "in expected_speakers list"}]
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "cli-expected"),
                                    lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "--expected-speakers", "carol,dave,erin", "-c", "podcast"])
    rec = saved(paths)
    assert rc == 0 and rec["context"] == "podcast" and [c["speaker_id"] for c in rec["mappings"]["A"]["candidates"]] == ["carol", "dave", "erin"]

    # the command line wins over the catalog for the speakers (speaker-assign:541-542), the catalog still names the context
    def both_sources(root):
        p = two(root, 2, "both")
        catalog_entry(root, p["audio"], "standup", ["alice"])
        return p
    rc, out, err, paths = pair.both(both_sources, lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "--expected-speakers", "zed"])
    rec = saved(paths)
    assert rec["context"] == "standup" and [c["speaker_id"] for c in rec["mappings"]["B"]["candidates"]] == ["zed"]


def test_assign_dry_run_text_and_json_save_nothing(pair):
    """:360, :394"""
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "dry"), lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "--dry-run",
                                                                         "--expected-speakers", "alice,bob"])
    assert rc == 0 and "=== DRY RUN - No changes saved ===" in out and "Assignments for: meeting.wav" in out
    assert not (paths["root"] / "assignments").exists() or not list((paths["root"] / "assignments").iterdir())
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "dryjson"), lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "-n", "-f", "json"])
    rec = json.loads(out[out.index("{"):])
    assert rc == 0 and rec["recording_b3sum"] == assign_cli.compute_b3sum(paths["audio"]) and rec["threshold"] == 0.3
    assert not (paths["root"] / "assignments").exists() or not list((paths["root"] / "assignments").iterdir())


def test_assign_saves_record_output_file_and_json(pair):
    """:437, :482, :712"""
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "save"),
                                    lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "-o", str(p["root"] / "result.yaml"), "--min-trust", "high"])
    rec = saved(paths)
    assert rc == 0 and "Assignments saved:" in out and rec["min_trust"] == "high" and rec["transcript_path"] == str(paths["t"].resolve())
    assert yaml.safe_load((paths["root"] / "result.yaml").read_text()) == rec
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "json"), lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "--format", "json"])
    rec = json.loads(out[out.index("{"):])
    assert rc == 0 and rec == json.loads(json.dumps(saved(paths)))


def test_assign_threshold_and_signal_combination(pair):
    """:761, :855 -- a context signal alone is worth 0.1: below the default threshold it only makes candidates, under a
    threshold of 0.05 the first expected speaker (insertion order decides ties: speaker-assign:461) is assigned"""
    argv = lambda thr: (lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "--expected-speakers", "alice,bob", "--threshold", thr, "-n", "-f", "json"])
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "thr-high"), argv("0.3"))
    rec = json.loads(out[out.index("{"):])
    assert rc == 0 and rec["mappings"]["A"]["speaker_id"] is None and [c["speaker_id"] for c in rec["mappings"]["A"]["candidates"]] == ["alice", "bob"]
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "thr-low"), argv("0.05"))
    rec = json.loads(out[out.index("{"):])
    assert rc == 0 and rec["mappings"]["A"]["speaker_id"] == "alice" and rec["mappings"]["A"]["confidence"] == "unassigned"
    assert rec["mappings"]["A"]["score"] == 0.1 and [c["speaker_id"] for c in rec["mappings"]["A"]["candidates"]] == ["bob"]


def test_assign_errors(pair):
    """:908, :932, :956"""
    rc, out, err, paths = pair.both(lambda r: {"audio": r / "nope.wav", "t": transcript(r / "transcript.json"), "root": r},
                                    lambda p: ["assign", str(p["audio"]), "-t", str(p["t"])])
    assert rc == 1 and "Error: Audio file not found:" in err
    rc, out, err, paths = pair.both(lambda r: {"audio": wav(r / "a.wav", "missing-t"), "t": r / "nope.json", "root": r},
                                    lambda p: ["assign", str(p["audio"]), "-t", str(p["t"])])
    assert rc == 1 and "Error: Transcript file not found:" in err

    def empty(root):
        t = root / "empty.json"
        t.write_text(json.dumps({"utterances": []}))
        return {"audio": wav(root / "a.wav", "empty"), "t": t, "root": root}
    rc, out, err, paths = pair.both(empty, lambda p: ["assign", str(p["audio"]), "-t", str(p["t"])])
    assert rc == 1 and "Error: No speakers found in transcript" in err


def test_assign_verbose_and_quiet(pair):
    """:989, :1017"""
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "verbose"),
                                    lambda p: ["-v", "assign", str(p["audio"]), "-t", str(p["t"]), "--expected-speakers", "alice"])
    assert rc == 0 and "Processing speaker A (3 segments)..." in out and "  Collecting context signals..." in out
    rc, out, err, paths = pair.both(lambda r: two(r, 2, "quiet"), lambda p: ["-q", "assign", str(p["audio"]), "-t", str(p["t"])])
    assert rc == 0 and out == "" and set(saved(paths)["mappings"]) == {"A", "B"}


def test_assign_speechmatics_transcript(pair):
    """:1049 -- labels S1 / S2 from results[].alternatives[].speaker; consecutive words of one speaker are one segment"""
    def sm(root):
        return {"audio": wav(root / "sm.wav", "speechmatics"), "t": transcript_speechmatics(root / "transcript_sm.json"), "root": root}
    rc, out, err, paths = pair.both(sm, lambda p: ["-v", "assign", str(p["audio"]), "-t", str(p["t"])])
    assert rc == 0 and "Found 2 speakers: S1, S2" in out and "Processing speaker S1 (2 segments)..." in out
    assert set(saved(paths)["mappings"]) == {"S1", "S2"}


def test_assign_with_the_embedding_step_through_a_stub_backend(pair, tmp_path, monkeypatch):
    """`assign -e` (speaker-assign:262-328, :556-570): the reference forks `speaker_detection identify` per label and keeps
    every row for every label; this repo calls the backend once, in process, and keeps whole-recording rows (no `label`
    key) for every label as well -- so with the same stub backend behind both (the reference's own `speaker_detection` on
    $PATH for its side) the records, the progress lines and the return codes are the same: min-trust filter (`low` rows
    dropped at `medium`, `unknown` kept), trust multipliers, default threshold and a lower one, text and JSON."""
    from speaker_diarization_toolkit_b200 import plugin_api
    from test_identify_cli_cpu import PROFILES, ROWS, STUB
    plug = tmp_path / "plug"
    plug.mkdir()
    (plug / "stub_backend_mod.py").write_text(STUB)
    cfg = tmp_path / "backends.yaml"
    cfg.write_text("backends:\n  stub:\n    module: stub_backend_mod\n")
    rows = tmp_path / "rows.json"
    rows.write_text(json.dumps({"rows": ROWS}))
    shim = tmp_path / "bin"
    shim.mkdir()
    (shim / "speaker_detection").write_text(f"#!/bin/sh\nexec {sys.executable} /root/reference/speaker_detection \"$@\"\n")
    (shim / "speaker_detection").chmod(0o755)
    root = Path(__file__).resolve().parent.parent
    pair.extra = {"SPEAKER_BACKENDS_CONFIG": str(cfg), "STUB_ROWS": str(rows), "SPEAKER_DETECTION_BACKEND": "stub",
                  "PYTHONPATH": os.pathsep.join([str(plug), str(root)]), "PATH": str(shim) + os.pathsep + os.environ.get("PATH", "")}
    monkeypatch.syspath_prepend(str(plug))
    monkeypatch.setattr(plugin_api, "_loaded", None)

    def build(uid):
        def go(r):
            (r / "db").mkdir(exist_ok=True)
            for pid, prof in PROFILES.items():
                (r / "db" / f"{pid}.json").write_text(json.dumps(dict(prof, version=1)))
            return two(r, 2, uid)
        return go

    rc, out, err, paths = pair.both(build("emb"), lambda p: ["-v", "assign", str(p["audio"]), "-t", str(p["t"]), "-e"])
    assert rc == 0 and err == "" and "  Collecting embedding signals..." in out and "    - alice: 0.91 (trust: low)" in out
    rec = saved(paths)
    for label in ("A", "B"):                                   # 0.4 x 0.4 x 0.91 = 0.1456 < 0.3: candidates only
        m = rec["mappings"][label]
        assert m["speaker_id"] is None and m["score"] == 0.146 and [c["speaker_id"] for c in m["candidates"]] == ["alice", "dave", "bob"]
        assert m["signals"] == [{"type": "embedding_match", "score": 0.91, "embedding_id": "emb-a2", "trust_level": "low", "backend": "stub"}]
    rc, out, err, paths = pair.both(build("emb-thr"), lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "-e", "--threshold", "0.1", "-f", "json"])
    rec = json.loads(out[out.index("{"):])
    assert rc == 0 and rec["mappings"]["A"]["speaker_id"] == "alice" and [c["speaker_id"] for c in rec["mappings"]["A"]["candidates"]] == ["dave", "bob"]
    rc, out, err, paths = pair.both(build("emb-trust"), lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "-e", "--min-trust", "medium",
                                                                  "--threshold", "0.1", "-n", "-f", "json"])
    rec = json.loads(out[out.index("{"):])                      # alice's row is `low`: dropped; `unknown` is not in the order: kept
    assert rc == 0 and rec["mappings"]["B"]["speaker_id"] == "dave" and rec["mappings"]["B"]["score"] == 0.124
    rc, out, err, paths = pair.both(build("emb-tags"), lambda p: ["assign", str(p["audio"]), "-t", str(p["t"]), "-e", "--tags", "team,eng",
                                                                 "--expected-speakers", "dave", "--threshold", "0.2", "-n", "-f", "json"])
    rec = json.loads(out[out.index("{"):])                      # bob is filtered by the tags; dave: 0.124 + 0.1 beats alice's 0.1456
    assert rc == 0 and rec["mappings"]["A"]["speaker_id"] == "dave" and rec["mappings"]["A"]["score"] == 0.224
    assert [s["type"] for s in rec["mappings"]["A"]["signals"]] == ["embedding_match", "context_expected"]
