"""Config 1 through the real command lines: synthetic store + Speechmatics-format transcript + sidecar ->
`speaker_detection identify --format json` and `speaker-assign assign --use-embeddings`, checked against the
C oracle (matching) and the restatement of the reference's combine_signals (assignment)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import matching_np as mnp
from speaker_diarization_toolkit_b200 import _native, store, synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def cli(script, *args, env):
    r = subprocess.run([sys.executable, str(ROOT / "bin" / script), *args], capture_output=True, text=True, env=env)
    return r.returncode, r.stdout, r.stderr


@pytest.fixture()
def cfg1_store(tmp_path):
    case = synth.config1()
    audio, tpath, ids = synth.write_store(case, tmp_path, ["S1", "S2"], speaker_names=["alice", "bob", "carol"])
    env = dict(os.environ, SPEAKERS_EMBEDDINGS_DIR=str(tmp_path), SPEAKER_DETECTION_BACKEND="b200", PYTHONPATH=str(ROOT))
    env.pop("SPEAKER_BACKENDS_CONFIG", None)
    return case, audio, tpath, ids, env


def expected_rows(oracle, case, ids, thr=0.354, k=10):
    rows, scores, counts = oracle.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=0, pool=0,
                                           threshold=thr, k=k)
    return rows, scores, counts


def test_identify_cli_config1(cfg1_store, oracle):
    case, audio, tpath, ids, env = cfg1_store
    rc, out, err = cli("speaker_detection", "identify", str(audio), "--format", "json", env=env)
    assert rc == 0, err
    got = json.loads(out)                                   # stdout is pure JSON, status went to stderr
    assert "Identifying speaker in rec.wav against 3 candidates..." in err
    rows, scores, counts = expected_rows(oracle, case, ids)
    exp = []
    trust = ["high", "medium", "low"]
    for g, label in enumerate(["S1", "S2"]):
        for i in range(counts[g]):
            r = int(rows[g, i])
            sid = ids[case.row_speaker[r]]
            exp.append({"speaker_id": sid, "name": sid.title(), "score": float(scores[g, i]), "confidence": float(scores[g, i]),
                        "trust_level": trust[case.row_trust[r]], "embedding_id": f"emb-{r:08x}", "backend": "b200",
                        "label": label, "rank": i})
    assert got == exp
    assert [r["speaker_id"] for r in got] == ["alice", "bob"]


def test_assign_cli_config1(cfg1_store, oracle):
    case, audio, tpath, ids, env = cfg1_store
    rc, out, err = cli("speaker-assign", "-q", "assign", str(audio), "-t", str(tpath), "--use-embeddings", "--format", "json",
                       "--threshold", "0.2", env=env)
    assert rc == 0, err
    got = json.loads(out)
    rows, scores, counts = expected_rows(oracle, case, ids)
    trust = ["high", "medium", "low"]
    for g, label in enumerate(["S1", "S2"]):
        sigs = [mnp.Signal("embedding_match", ids[case.row_speaker[int(rows[g, i])]], float(scores[g, i]),
                           {"embedding_id": f"emb-{int(rows[g, i]):08x}", "trust_level": trust[case.row_trust[int(rows[g, i])]],
                            "backend": "b200"}) for i in range(counts[g])]
        ref = mnp.combine_signals(label, sigs, threshold=0.2)
        m = got["mappings"][label]
        assert m["speaker_id"] == ref["speaker_id"]
        assert m["confidence"] == ref["confidence"]
        assert m["score"] == round(ref["score"], 3)
        assert m["signals"] == ref["signals"]
        assert m.get("candidates", []) == ref["candidates"]
    # alice is enrolled at high trust (0.4*1.0*0.89 = 0.36 -> assigned); bob at medium trust (0.4*0.7*0.89 = 0.25 -> assigned at 0.2)
    assert got["mappings"]["S1"]["speaker_id"] == "alice" and got["mappings"]["S2"]["speaker_id"] == "bob"
    saved = list((Path(env["SPEAKERS_EMBEDDINGS_DIR"]) / "assignments").glob("*.yaml"))
    assert len(saved) == 1 and saved[0].stem == got["recording_b3sum"]


def test_assign_cli_default_threshold_needs_high_trust(cfg1_store):
    """With speaker-assign's default 0.3 an embedding-only match is impossible at medium trust."""
    case, audio, tpath, ids, env = cfg1_store
    rc, out, err = cli("speaker-assign", "-q", "assign", str(audio), "-t", str(tpath), "-e", "-n", "--format", "json", env=env)
    assert rc == 0, err
    got = json.loads(out[out.index("{"):])
    assert got["mappings"]["S1"]["speaker_id"] == "alice"
    assert got["mappings"]["S2"]["speaker_id"] is None and got["mappings"]["S2"]["candidates"][0]["speaker_id"] == "bob"


def test_verify_cli(cfg1_store):
    case, audio, tpath, ids, env = cfg1_store
    rc, out, err = cli("speaker_detection", "verify", "carol", str(audio), env=env)
    assert rc == 1 and "NO MATCH" in out                    # carol speaks in neither label
    # a recording of alice only (the S1 segments of config 1)
    from speaker_diarization_toolkit_b200 import store
    a2 = Path(env["SPEAKERS_EMBEDDINGS_DIR"]) / "solo.wav"
    a2.write_bytes(b"RIFFsolo")
    solo = case.seg[case.seg_label == 0]
    store.save_segment_embeddings(a2, "b200", solo, ["S1"] * len(solo))
    rc, out, err = cli("speaker_detection", "verify", "alice", str(a2), env=env)
    assert rc == 0 and "MATCH: Speaker 'alice' verified (confidence: 0." in out


def build_batch_store(tmp_path, n_rec=5):
    """A store with 12 enrolled speakers (17 bank rows) and n_rec recordings; returns (manifest, manifest path, env)."""
    bank_case = synth.make_case(7, [1], 12, 192, rows_per_speaker=[1, 2, 1, 3, 1, 1, 2, 1, 1, 1, 2, 1], trust_cycle=(0, 0, 1, 2))
    ids = [f"spk{idx:04d}" for idx in range(12)]
    manifest = []
    env = dict(os.environ, SPEAKERS_EMBEDDINGS_DIR=str(tmp_path), SPEAKER_DETECTION_BACKEND="b200", PYTHONPATH=str(ROOT))
    env.pop("SPEAKER_BACKENDS_CONFIG", None)
    from speaker_diarization_toolkit_b200.synth import Case
    rng = np.random.default_rng(3)
    for r in range(n_rec):
        labels = [f"S{i + 1}" for i in range(2 + r % 3)]
        counts = rng.integers(3, 30, size=len(labels))
        rec = synth.make_case(100 + r, counts, 12, 192, truth=list(rng.integers(0, 12, size=len(labels))))
        # same enrolled speakers for every recording: regenerate the segments around the bank's centroids
        merged = Case(rec.seg, rec.seg_label, rec.goff, bank_case.bank, bank_case.row_speaker, bank_case.row_trust, rec.truth, 12)
        for g in range(len(labels)):
            spk = int(rec.truth[g])
            row = int(np.flatnonzero(bank_case.row_speaker == spk)[0])
            base = bank_case.bank[row] / np.linalg.norm(bank_case.bank[row])
            n = int(counts[g])
            x = base[None, :] + 0.25 / np.sqrt(192) * rng.standard_normal((n, 192))
            merged.seg[rec.goff[g]:rec.goff[g + 1]] = x.astype(np.float32)
        audio, tpath, _ = synth.write_store(merged, tmp_path, labels, speaker_names=ids, audio_name=f"rec{r}.wav")
        manifest.append({"audio": str(audio), "transcript": str(tpath)})
    mpath = tmp_path / "manifest.json"
    mpath.write_text(json.dumps(manifest))
    return manifest, mpath, env


def strip_time(outputs):
    return [{k: v for k, v in o.items() if k != "assigned_at"} for o in outputs]


def test_assign_batch_equals_per_recording_assign(tmp_path, oracle):
    """`assign-batch` (one resident backend call for all recordings, assignment on the device) writes
    the same mappings as one `assign --use-embeddings` per recording."""
    manifest, mpath, env = build_batch_store(tmp_path)
    singles = {}
    for item in manifest:
        rc, out, err = cli("speaker-assign", "-q", "assign", item["audio"], "-t", item["transcript"], "-e", "-n", "--format", "json",
                           "--threshold", "0.2", env=env)
        assert rc == 0, err
        singles[item["audio"]] = json.loads(out[out.index("{"):])["mappings"]
    rc, out, err = cli("speaker-assign", "-q", "assign-batch", str(mpath), "--threshold", "0.2", "--format", "json", env=env)
    assert rc == 0, err
    batch = json.loads(out)
    assert len(batch) == 5
    n_assigned = 0
    for item, b in zip(manifest, batch):
        assert b["mappings"] == singles[item["audio"]]
        n_assigned += sum(1 for m in b["mappings"].values() if m["speaker_id"])
    assert n_assigned >= 5
    assert len(list((tmp_path / "assignments").glob("*.yaml"))) == 5
    # `--gpus 3`: one worker process per GPU, the manifest split over them (workers share the device on a 1-GPU box);
    # same output, same order, every recording written
    for f in (tmp_path / "assignments").glob("*.yaml"):
        f.unlink()
    rc, out, err = cli("speaker-assign", "-q", "assign-batch", str(mpath), "--threshold", "0.2", "--format", "json", "--gpus", "3", env=env)
    assert rc == 0, err
    assert strip_time(json.loads(out)) == strip_time(batch)
    assert len(list((tmp_path / "assignments").glob("*.yaml"))) == 5


def test_review_consistency_flags_mislabelled_segments(tmp_path, oracle):
    """the config-5 pooled affinity behind `speaker-review consistency`: segments planted under the
    wrong diarization label are exactly the ones reported, with the affinities the oracle computes."""
    case = synth.make_case(21, [300, 260, 280, 40], 4, 256, truth=[0, 1, 2, 3], impostor_frac=0.0)
    lab = case.seg_label.copy()
    wrong = [5, 17, 400, 610, 845]                      # move five segments under a neighbouring label
    for i in wrong:
        lab[i] = (lab[i] + 1) % 4
    labels = [f"S{int(l) + 1}" for l in lab]
    audio = tmp_path / "m.wav"
    audio.write_bytes(b"RIFF" + bytes(40) + b"review")
    start = np.arange(len(lab)) * 1.0
    store.save_segment_embeddings(audio, "b200", case.seg, labels, start, start + 0.9)
    env = dict(os.environ, PYTHONPATH=str(ROOT), SPEAKER_DETECTION_BACKEND="b200")
    rc, out, err = cli("speaker-review", "consistency", str(audio), "--format", "json", "--fail-on-suspects", env=env)
    assert rc == 2, err
    rep = json.loads(out)
    se = store.load_segment_embeddings(audio, "b200")          # sidecar order: stable sort by label
    flagged = sorted(round(s["start"]) for s in rep["suspects"])
    assert flagged == sorted(wrong)
    for s in rep["suspects"]:
        assert s["closer_to"] == f"S{int(case.seg_label[round(s['start'])]) + 1}"       # points back at the true label
    goff = np.r_[0, np.cumsum(np.bincount(se.label_index, minlength=4))].astype(np.int64)
    ref = oracle.affinity(se.emb, goff, mode=1, pool=0)
    for s in rep["suspects"]:
        i = s["segment"]
        assert abs(s["affinity"] - ref[i, se.label_index[i]]) < 2e-5
    assert np.allclose(np.diag(np.asarray(rep["label_affinity"])), [np.mean(ref[goff[a]:goff[a + 1], a]) for a in range(4)], atol=2e-5)
    rc, out, err = cli("speaker-review", "consistency", str(audio), env=env)
    assert rc == 0 and "5 suspect segment(s)" in out
