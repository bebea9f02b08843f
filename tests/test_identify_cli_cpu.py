"""`speaker_detection identify | verify` (speaker_detection:1031-1178) side by side with the REFERENCE CLI, a stub backend
plugged into both through the plugin registry ($SPEAKER_BACKENDS_CONFIG -> module -> Backend(), base.py:212-293): what is
compared is everything the CLI does around the backend call -- tag filter, candidate selection, the status line on stderr,
decoration of the rows with name / trust level / embedding id, both output formats, verify's MATCH / NO MATCH lines, error
strings and return codes.  Runs where /root/reference is present (this container); without it the assertions on this
repo's own output still run.  No device is touched: the stub stands where the B200 backend would."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from speaker_diarization_toolkit_b200 import identify_cli, plugin_api

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/speaker_detection")

STUB = '''
import json, os
from pathlib import Path
try:
    from speaker_detection_backends.base import EmbeddingBackend        # under the reference CLI
except ImportError:
    from speaker_diarization_toolkit_b200.plugin_api import EmbeddingBackend


class Backend(EmbeddingBackend):
    """answers from the JSON file named by $STUB_ROWS: {"rows": [...]} or {"raise": "message"}"""
    @property
    def name(self):
        return "stub"

    @property
    def requires_api_key(self):
        return False

    @property
    def model_version(self):
        return "stub-v1"

    def enroll_speaker(self, audio_path, segments=None):
        return {"external_id": "x", "model_version": self.model_version}

    def identify_speaker(self, audio_path, candidates, threshold=0.354):
        spec = json.loads(Path(os.environ["STUB_ROWS"]).read_text())
        if "raise" in spec:
            raise RuntimeError(spec["raise"])
        ids = {c["id"] for c in candidates}
        return [dict(r) for r in spec["rows"] if r["speaker_id"] in ids and r.get("confidence", r.get("similarity", 1.0)) >= threshold]

    def verify_speaker(self, audio_path, speaker_profile, threshold=0.354):
        rows = self.identify_speaker(audio_path, [speaker_profile], threshold)
        if rows:
            s = rows[0].get("confidence", rows[0].get("similarity"))
            return {"match": True, "similarity": s, "confidence": s, "embedding_id": rows[0].get("embedding_id")}
        return {"match": False, "similarity": 0.0, "confidence": 0.0, "embedding_id": None}
'''

PROFILES = {
    "alice": {"id": "alice", "names": {"default": "Alice Anderson"}, "tags": ["team", "eng"],
              "embeddings": {"stub": [{"id": "emb-a1", "trust_level": "high"}, {"id": "emb-a2", "trust_level": "low"}]}},
    "bob": {"id": "bob", "names": {"default": "Bob"}, "tags": ["team"],
            "embeddings": {"stub": [{"id": "emb-b1", "trust_level": "medium"}], "other": [{"id": "o1"}]}},
    "carol": {"id": "carol", "names": {"default": "Carol"}, "tags": ["guest"], "embeddings": {"other": [{"id": "o2"}]}},
    "dave": {"id": "dave", "names": {"default": "Dave"}, "tags": ["team", "eng"],
             "embeddings": {"stub": [{"id": "emb-d1"}, {"id": "emb-d2", "trust_level": "invalidated"}]}},
}
ROWS = [
    {"speaker_id": "alice", "similarity": 0.91, "embedding_id": "emb-a2"},        # trust looked up by embedding id
    {"speaker_id": "dave", "similarity": 0.62},                                    # no id: best trust level of the speaker's embeddings
    {"speaker_id": "bob", "confidence": 0.4, "similarity": 0.1, "embedding_id": "nope"},     # `confidence` wins; unknown id -> unknown trust
]


@pytest.fixture()
def env(tmp_path, monkeypatch):
    (tmp_path / "db").mkdir()
    for pid, p in PROFILES.items():
        (tmp_path / "db" / f"{pid}.json").write_text(json.dumps(dict(p, version=1)))      # current schema: the reference migrates older files in place
    mod = tmp_path / "plug"
    mod.mkdir()
    (mod / "stub_backend_mod.py").write_text(STUB)
    cfg = tmp_path / "backends.yaml"
    cfg.write_text("backends:\n  stub:\n    module: stub_backend_mod\n")
    rows = tmp_path / "rows.json"
    rows.write_text(json.dumps({"rows": ROWS}))
    audio = tmp_path / "clip.wav"
    audio.write_bytes(b"RIFF0000WAVEfmt ")
    e = {"SPEAKERS_EMBEDDINGS_DIR": str(tmp_path), "SPEAKER_BACKENDS_CONFIG": str(cfg), "STUB_ROWS": str(rows)}
    for k, v in e.items():
        monkeypatch.setenv(k, v)
    monkeypatch.syspath_prepend(str(mod))
    monkeypatch.setattr(plugin_api, "_loaded", None)     # the registry caches its config per process (restored afterwards)
    return {"root": tmp_path, "audio": audio, "rows": rows, "env": e, "mod": mod}


def both(env, capsys, argv):
    rc = identify_cli.main(argv)
    cap = capsys.readouterr()
    if REF.exists():
        r = subprocess.run([sys.executable, str(REF), *argv], capture_output=True, text=True,
                           env=dict(os.environ, **env["env"], PYTHONPATH=os.pathsep.join([str(env["mod"]), str(ROOT)])))
        assert (r.returncode, r.stdout, r.stderr) == (rc, cap.out, cap.err), argv
    return rc, cap.out, cap.err


def test_identify_json_and_text(env, capsys):
    a = str(env["audio"])
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub", "--format", "json"])
    got = json.loads(out)
    assert rc == 0 and err == "Identifying speaker in clip.wav against 3 candidates...\n"
    assert [(g["speaker_id"], g["name"], g["score"], g["trust_level"], g["embedding_id"], g["backend"]) for g in got] == [
        ("alice", "Alice Anderson", 0.91, "low", "emb-a2", "stub"), ("dave", "Dave", 0.62, "unknown", None, "stub"),
        ("bob", "Bob", 0.4, "unknown", "nope", "stub")]
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub"])
    assert rc == 0 and out == "\nMatches:\n  alice: Alice Anderson (confidence: 0.91)\n  dave: Dave (confidence: 0.62)\n  bob: Bob (confidence: 0.40)\n"


def test_identify_tags_threshold_and_backend_from_the_environment(env, capsys, monkeypatch):
    a = str(env["audio"])
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub", "--tags", "team,eng", "-f", "json"])        # all tags must match
    assert rc == 0 and "against 2 candidates" in err and [g["speaker_id"] for g in json.loads(out)] == ["alice", "dave"]
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub", "--threshold", "0.7", "-f", "json"])
    assert rc == 0 and [g["speaker_id"] for g in json.loads(out)] == ["alice"]
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub", "--threshold", "0.95", "-f", "json"])
    assert rc == 0 and out == "[]\n"
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub", "--threshold", "0.95"])
    assert rc == 0 and out == "No matching speakers found.\n"
    monkeypatch.setenv("SPEAKER_DETECTION_BACKEND", "stub")
    env["env"]["SPEAKER_DETECTION_BACKEND"] = "stub"
    rc, out, err = both(env, capsys, ["identify", a, "-f", "json"])
    assert rc == 0 and len(json.loads(out)) == 3


def test_identify_errors(env, capsys):
    a = str(env["audio"])
    rc, out, err = both(env, capsys, ["identify", str(env["root"] / "missing.wav"), "-b", "stub"])
    assert rc == 1 and err.startswith("Error: Audio file not found:")
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub", "--tags", "nobody"])
    assert rc == 1 and err == "No speakers to match against.\n"
    rc, out, err = both(env, capsys, ["identify", a, "-b", "third"])
    assert rc == 1 and err == "No speakers with third embeddings.\n"
    rc, out, err = both(env, capsys, ["identify", a, "-b", "other"])
    assert rc == 1 and err.startswith("Error loading backend: Unknown backend: other. Available: stub")
    env["rows"].write_text(json.dumps({"raise": "the extractor is down"}))
    rc, out, err = both(env, capsys, ["identify", a, "-b", "stub"])
    assert rc == 1 and err.endswith("Error during identification: the extractor is down\n") and out == ""


def test_verify(env, capsys):
    a = str(env["audio"])
    rc, out, err = both(env, capsys, ["verify", "alice", a, "-b", "stub"])
    assert rc == 0 and out == "MATCH: Speaker 'alice' verified (confidence: 0.91)\n" and err == "Verifying audio against speaker 'alice'...\n"
    rc, out, err = both(env, capsys, ["verify", "alice", a, "-b", "stub", "--threshold", "0.95"])
    assert rc == 1 and out == "NO MATCH: Audio does not match speaker 'alice'\n"
    rc, out, err = both(env, capsys, ["verify", "nobody", a, "-b", "stub"])
    assert rc == 1 and err == "Error: Speaker 'nobody' not found.\n"
    rc, out, err = both(env, capsys, ["verify", "carol", a, "-b", "stub"])
    assert rc == 1 and err == "Error: Speaker 'carol' has no stub embeddings.\n"
    rc, out, err = both(env, capsys, ["verify", "alice", str(env["root"] / "missing.wav"), "-b", "stub"])
    assert rc == 1 and err.startswith("Error: Audio file not found:")
    env["rows"].write_text(json.dumps({"raise": "boom"}))
    rc, out, err = both(env, capsys, ["verify", "alice", a, "-b", "stub"])
    assert rc == 1 and err.endswith("Error during verification: boom\n")
