"""CPU tests of the batch glue (batch.py, `speaker-assign assign-batch`) with the device replaced by a
stand-in context that answers from the oracle.  What is under test is the HOST logic: concatenation of the recordings
into label groups, the mapping of bank rows / label groups back to speaker ids and diarization labels, the
embedding-only (device assignment) and context-signal (Python fusion) branches, and the YAML / JSON output of the
command.  The arithmetic itself is covered by the GPU parity tests."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import canonical
from speaker_diarization_toolkit_b200 import _native, assign_cli, batch, review_cli, signals as sg, store, synth


class OracleContext:
    """Same surface as _native.Context (bank_load / identify / assign / fetch / close), answers from oracle/canonical."""

    def __init__(self, *a, **k):
        self.closed = False

    def close(self):
        self.closed = True

    def bank_load(self, rows, row_speaker, row_trust=None, dtype=_native.DTYPE_F32, global_row_offset=0):
        self.bank, self.spk, self.dtype = np.asarray(rows, np.float32), np.asarray(row_speaker, np.int32), dtype
        self.trust = np.asarray(row_trust, np.uint8) if row_trust is not None else np.full(len(self.spk), 4, np.uint8)

    def identify(self, seg, seg_label, L, pool=0, threshold=0.354, k=10):
        lab = np.asarray(seg_label, np.int64)
        goff = np.searchsorted(lab, np.arange(L + 1)).astype(np.int64)
        self.rows, self.scores, self.counts = canonical.identify(seg, goff, self.bank, self.spk, int(self.spk.max()) + 1,
                                                                 mode=self.dtype, pool=pool, threshold=threshold, k=k)
        self.k = k
        return self.rows, self.scores, self.counts

    def assign(self, assign_threshold=0.3, min_trust="low"):
        t = np.full(self.rows.shape, 4, np.uint8)
        ok = self.rows >= 0
        t[ok] = self.trust[self.rows[ok]]
        self.t = t
        self.a = canonical.assign(self.rows, self.scores, t, self.counts, assign_threshold, _native.TRUST_CODES.get(min_trust, 99))

    def affinity_pooled(self, seg, seg_label, L, dtype=_native.DTYPE_BF16, pool=0, want_ll=True):
        lab = np.asarray(seg_label, np.int64)
        goff = np.searchsorted(lab, np.arange(L + 1)).astype(np.int64)
        nl = canonical.affinity(seg, goff, mode=dtype, pool=pool)
        q = np.rint(nl.astype(np.float64) * 2.0 ** 30)
        ll = np.stack([q[goff[a]:goff[a + 1]].sum(axis=0) / (max(1, goff[a + 1] - goff[a]) * 2.0 ** 30) for a in range(L)]).astype(np.float32)
        return nl, ll

    def fetch(self, with_assign=False):
        out = {"row": self.rows, "score": self.scores, "count": self.counts, "trust": self.t}
        if with_assign:
            out.update(dict(zip(["assign_idx", "assign_score", "assign_conf", "cand_idx", "cand_score"], self.a)))
        return out


@pytest.fixture()
def store_with_recordings(tmp_path, monkeypatch):
    monkeypatch.setenv("SPEAKERS_EMBEDDINGS_DIR", str(tmp_path))
    monkeypatch.setattr(_native, "Context", OracleContext)
    bank_case = synth.make_case(7, [1], 9, 48, rows_per_speaker=[1, 2, 1, 3, 1, 1, 2, 1, 1], trust_cycle=(0, 0, 1, 2))
    ids = [f"spk{idx:02d}" for idx in range(9)]
    rng = np.random.default_rng(3)
    manifest = []
    for r in range(4):
        labels = [f"S{i + 1}" for i in range(1 + r)]
        counts = rng.integers(2, 12, size=len(labels))
        truth = list(rng.integers(0, 9, size=len(labels)))
        rec = synth.make_case(100 + r, counts, 9, 48, truth=truth)
        merged = synth.Case(rec.seg, rec.seg_label, rec.goff, bank_case.bank, bank_case.row_speaker, bank_case.row_trust, rec.truth, 9)
        for g, spk in enumerate(truth):
            row = int(np.flatnonzero(bank_case.row_speaker == spk)[0])
            base = bank_case.bank[row] / np.linalg.norm(bank_case.bank[row])
            n = int(counts[g])
            merged.seg[rec.goff[g]:rec.goff[g + 1]] = (base[None, :] + 0.2 / np.sqrt(48) * rng.standard_normal((n, 48))).astype(np.float32)
        audio, tpath, _ = synth.write_store(merged, tmp_path, labels, speaker_names=ids, audio_name=f"rec{r}.wav")
        manifest.append({"audio": str(audio), "transcript": str(tpath), "truth": [ids[t] for t in truth], "labels": labels})
    return tmp_path, manifest, ids


def test_batch_matcher_equals_one_recording_at_a_time(store_with_recordings):
    root, manifest, ids = store_with_recordings
    m = batch.BatchMatcher(threshold=0.354)
    m.load_bank()
    together = m.identify([it["audio"] for it in manifest], assign_threshold=0.2, min_trust="low")
    assert [r.labels for r in together] == [it["labels"] for it in manifest]
    for it, res in zip(manifest, together):
        alone = m.identify([it["audio"]], assign_threshold=0.2, min_trust="low")[0]
        assert res.mappings == alone.mappings and res.matches == alone.matches      # label groups do not leak across recordings
        for label, want in zip(it["labels"], it["truth"]):
            rows = res.matches[label]
            assert rows and rows[0]["speaker_id"] == want and rows[0]["label"] == label and rows[0]["rank"] == 0
            assert [r["score"] for r in rows] == sorted((r["score"] for r in rows), reverse=True)
            # the device-side assignment agrees with the Python restatement of combine_signals on the same rows
            ref = sg.combine_signals(label, sg.signals_from_matches(rows, min_trust="low", label=label), threshold=0.2)
            got = res.mappings[label]
            assert got["speaker_id"] == ref.speaker_id and got["confidence"] == ref.confidence and got["score"] == round(ref.score, 3)
            assert got.get("candidates", []) == ref.candidates
    m.close()
    assert m.ctx.closed


def test_batch_matcher_context_signals_take_the_python_fusion(store_with_recordings):
    root, manifest, ids = store_with_recordings
    m = batch.BatchMatcher(threshold=0.354)
    m.load_bank()
    it = manifest[2]
    other = next(i for i in ids if i not in it["truth"])
    res = m.identify([it["audio"]], assign_threshold=0.2, min_trust="low", expected=[("standup", [it["truth"][0], other])])[0]
    m0 = res.mappings[it["labels"][0]]
    types = sorted(s["type"] for s in m0["signals"])
    assert "context_expected" in types and "embedding_match" in types        # speaker-assign:331-356 + :262-328 fused
    assert m0["speaker_id"] == it["truth"][0]


def test_assign_batch_command(store_with_recordings, capsys):
    root, manifest, ids = store_with_recordings
    mpath = root / "manifest.jsonl"
    mpath.write_text("\n".join(json.dumps({"audio": it["audio"], "transcript": it["transcript"]}) for it in manifest))
    rc = assign_cli.main(["-q", "assign-batch", str(mpath), "--threshold", "0.2", "--format", "json"])
    out = json.loads(capsys.readouterr().out)
    assert rc == 0 and len(out) == 4
    n_assigned = 0
    for it, o in zip(manifest, out):
        assert list(o["mappings"]) == it["labels"] and o["threshold"] == 0.2 and o["min_trust"] == "low"
        for label, want in zip(it["labels"], it["truth"]):
            got = o["mappings"][label]
            # a low-trust enrolment cannot reach the assignment threshold (0.4 * 0.4 * cosine < 0.2): then the true speaker
            # is the first candidate instead
            assert got["speaker_id"] in (want, None)
            if got["speaker_id"] is None:
                assert got["confidence"] == "unassigned" and got["candidates"][0]["speaker_id"] == want
            n_assigned += got["speaker_id"] is not None
    assert n_assigned >= 5
    assert len(list((root / "assignments").glob("*.yaml"))) == 4
    # a missing file is reported with the reference's message and nothing is written
    bad = root / "bad.json"
    bad.write_text(json.dumps([{"audio": str(root / "nope.wav"), "transcript": manifest[0]["transcript"]}]))
    assert assign_cli.main(["assign-batch", str(bad)]) == 1
    assert "Error: Audio file not found:" in capsys.readouterr().err


def test_review_consistency_host_logic(tmp_path, monkeypatch, capsys):
    """`speaker-review consistency`: which segments are reported, how they are ordered, exit codes."""
    monkeypatch.setattr(_native, "Context", OracleContext)
    case = synth.make_case(21, [60, 50, 40], 3, 32, truth=[0, 1, 2], impostor_frac=0.0)
    lab = case.seg_label.copy()
    wrong = [3, 70, 120]
    for i in wrong:
        lab[i] = (lab[i] + 1) % 3
    audio = tmp_path / "m.wav"
    audio.write_bytes(b"RIFF" + bytes(40))
    start = np.arange(len(lab)) * 1.0
    store.save_segment_embeddings(audio, "b200", case.seg, [f"S{int(l) + 1}" for l in lab], start, start + 0.9)
    rep = review_cli.consistency(audio)
    assert sorted(round(s["start"]) for s in rep["suspects"]) == wrong and rep["segments_checked"] == 150
    assert [s["gap"] for s in rep["suspects"]] == sorted(s["gap"] for s in rep["suspects"])       # worst first
    for s in rep["suspects"]:
        assert s["closer_to"] == f"S{int(case.seg_label[round(s['start'])]) + 1}" and s["closer_affinity"] > s["affinity"]
    assert np.asarray(rep["label_affinity"]).shape == (3, 3)
    assert review_cli.main(["consistency", str(audio), "--fail-on-suspects"]) == 2
    assert "3 suspect segment(s)" in capsys.readouterr().out
    assert review_cli.main(["consistency", str(audio), "--margin", "-1.0", "--format", "json", "--fail-on-suspects"]) == 0
    assert json.loads(capsys.readouterr().out)["suspects"] == []
    assert review_cli.main(["consistency", str(tmp_path / "nope.wav")]) == 1
    assert "Error: Audio file not found:" in capsys.readouterr().err
