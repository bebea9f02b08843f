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


# ---- multi-GPU orchestration of assign-batch, host logic on CPU -----------------------------------------
class _InProcessWorker:
    """Stand-in for subprocess.Popen: runs the worker command line of `assign-batch --gpus N` in this process (with the
    oracle-backed context), so the partition of the manifest and the merge of the workers' outputs are tested on CPU."""
    launched = []
    real = None

    def __new__(cls, cmd, *a, **k):
        if list(cmd[1:3]) != ["-m", "speaker_diarization_toolkit_b200.assign_cli"]:
            return cls.real(cmd, *a, **k)            # anything else (b3sum ...) is a real subprocess
        return super().__new__(cls)

    def __init__(self, cmd, env=None, stdout=None, stderr=None, text=None):
        import contextlib
        import io
        import os
        assert cmd[1:4] == ["-m", "speaker_diarization_toolkit_b200.assign_cli", "assign-batch"]
        _InProcessWorker.launched.append((cmd, {k: v for k, v in env.items() if k.startswith("SPEAKER_B200_")}))
        out, err = io.StringIO(), io.StringIO()
        saved = {k: os.environ.get(k) for k in ("SPEAKER_B200_DEVICE",)}
        os.environ["SPEAKER_B200_DEVICE"] = env["SPEAKER_B200_DEVICE"]
        try:
            with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
                self.returncode = assign_cli.main(cmd[3:])
        finally:
            for k, v in saved.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        # (NCCL writes its version line to stdout under NCCL_DEBUG=VERSION: the parent must cope with such noise)
        self._out, self._err = "NCCL version 2.27.3+cuda12.9\n" + out.getvalue(), err.getvalue()

    def communicate(self):
        return self._out, self._err


def test_assign_batch_gpus_splits_recordings_and_concatenates(store_with_recordings, capsys, monkeypatch):
    import subprocess
    root, manifest, ids = store_with_recordings
    mpath = root / "manifest.json"
    mpath.write_text(json.dumps([{"audio": it["audio"], "transcript": it["transcript"]} for it in manifest]))
    assert assign_cli.main(["-q", "assign-batch", str(mpath), "--threshold", "0.2", "--format", "json", "--dry-run"]) == 0
    single = json.loads(capsys.readouterr().out)
    _InProcessWorker.real, _InProcessWorker.launched = subprocess.Popen, []
    monkeypatch.setattr(subprocess, "Popen", _InProcessWorker)
    monkeypatch.setattr(_native, "device_count", lambda: 3)
    assert assign_cli.main(["-q", "assign-batch", str(mpath), "--threshold", "0.2", "--format", "json", "--gpus", "3"]) == 0
    multi = json.loads(capsys.readouterr().out)
    strip = lambda o: {k: v for k, v in o.items() if k != "assigned_at"}
    assert [strip(o) for o in multi] == [strip(o) for o in single]              # same answers, same order
    devs = [env["SPEAKER_B200_DEVICE"] for _, env in _InProcessWorker.launched]
    assert len(devs) >= 2 and devs == sorted(set(devs)) and all("SPEAKER_B200_WORLD" not in env for _, env in _InProcessWorker.launched)
    assert len(list((root / "assignments").glob("*.yaml"))) == 4                # every worker wrote its recordings


class _ShardContext(OracleContext):
    """Two of these (world = 2) emulate the row-sharded library: local top-k on the shard, an in-process 'all-gather'
    of the lists, the host mirror of the merge kernel."""
    pool = {}

    def __init__(self, device=0, world=1, rank=0, uid=None):
        super().__init__()
        self.world, self.rank = world, rank

    def bank_load(self, rows, row_speaker, row_trust=None, dtype=_native.DTYPE_F32, global_row_offset=0):
        super().bank_load(rows, row_speaker, row_trust, dtype, global_row_offset)
        self.off = global_row_offset

    def identify(self, seg, seg_label, L, pool=0, threshold=0.354, k=10):
        import threading
        from speaker_diarization_toolkit_b200 import sharding
        lab = np.asarray(seg_label, np.int64)
        goff = np.searchsorted(lab, np.arange(L + 1)).astype(np.int64)
        if len(self.spk):
            base = int(self.spk.min())
            r, s, c = canonical.identify(np.asarray(seg, np.float32), goff, self.bank, self.spk - base, int(self.spk.max()) - base + 1,
                                         mode=self.dtype, pool=pool, threshold=threshold, k=k)
            r = np.where(r >= 0, r + self.off, r)
        else:
            r, s, c = np.full((L, k), -1, np.int64), np.zeros((L, k), np.float32), np.zeros(L, np.int32)
        _ShardContext.pool[self.rank] = (r, s, c, self.trust, self.off)
        _ShardContext.barrier.wait(timeout=60)
        parts = [_ShardContext.pool[i] for i in range(self.world)]
        self.rows, self.scores, self.counts, _ = sharding.merge_topk_lists(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]),
                                                                           np.stack([p[2] for p in parts]), k)
        self.gtrust = np.concatenate([p[3] for p in parts])
        self.k = k
        _ShardContext.barrier.wait(timeout=60)
        return self.rows, self.scores, self.counts

    def assign(self, assign_threshold=0.3, min_trust="low"):
        t = np.full(self.rows.shape, 4, np.uint8)
        ok = self.rows >= 0
        t[ok] = self.gtrust[self.rows[ok]]
        self.t = t
        self.a = canonical.assign(self.rows, self.scores, t, self.counts, assign_threshold, _native.TRUST_CODES.get(min_trust, 99))


def test_batch_matcher_row_sharded_world2_equals_single(store_with_recordings, monkeypatch):
    import threading
    root, manifest, ids = store_with_recordings
    audios = [it["audio"] for it in manifest]
    m1 = batch.BatchMatcher(threshold=0.354)
    m1.load_bank()
    want = m1.identify(audios, assign_threshold=0.2, min_trust="low")
    monkeypatch.setattr(_native, "Context", _ShardContext)
    _ShardContext.barrier, _ShardContext.pool = threading.Barrier(2), {}
    got, errs = {}, []

    def run(rank):
        try:
            m = batch.BatchMatcher(threshold=0.354, world=2, rank=rank, nccl_uid=b"x" * 128)
            bank = m.load_bank()
            assert m.ctx.off == ([0] + [p1 for _, p1 in __import__("speaker_diarization_toolkit_b200").sharding.shard_bank_rows(bank.row_speaker, 2)])[rank]
            got[rank] = m.identify(audios, assign_threshold=0.2, min_trust="low")
        except Exception as exc:       # pragma: no cover
            errs.append(repr(exc))
            _ShardContext.barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert not errs, errs
    for rank in range(2):                       # every rank ends up with the global answer
        assert [(r.labels, r.matches, r.mappings) for r in got[rank]] == [(r.labels, r.matches, r.mappings) for r in want]


def test_sharded_context_from_env_publishes_the_unique_id(tmp_path, monkeypatch):
    made = []
    monkeypatch.setattr(_native, "Context", lambda *a: made.append(a) or "ctx")
    monkeypatch.setattr(_native, "nccl_unique_id", lambda: bytes(range(128)))
    assert batch.sharded_context_from_env(3)[:2] == (1, 0) and made[-1] == (3,)
    monkeypatch.setenv("SPEAKER_B200_WORLD", "2")
    with pytest.raises(ValueError):
        batch.sharded_context_from_env(0)
    monkeypatch.setenv("SPEAKER_B200_UID_FILE", str(tmp_path / "uid"))
    monkeypatch.setenv("SPEAKER_B200_RANK", "0")
    assert batch.sharded_context_from_env(0)[:2] == (2, 0) and made[-1] == (0, 2, 0, bytes(range(128)))
    monkeypatch.setenv("SPEAKER_B200_RANK", "1")
    assert batch.sharded_context_from_env(1)[:2] == (2, 1) and made[-1] == (1, 2, 1, bytes(range(128)))   # reads what rank 0 wrote
    (tmp_path / "uid").unlink()
    monkeypatch.setenv("SPEAKER_B200_UID_TIMEOUT", "0.1")
    with pytest.raises(TimeoutError):
        batch.sharded_context_from_env(1)


def test_fp16_sidecar_and_long_labels_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    emb = rng.standard_normal((6, 32)).astype(np.float32)
    long_label = "speaker-" + "x" * 60                      # no 32-character cut (the reference has no label length limit)
    labels = [long_label, "S1", long_label, "S1", "S2", "S2"]
    audio = tmp_path / "a.wav"
    store.save_segment_embeddings(audio, "b200", emb, labels, dtype=np.float16)
    assert store.sidecar_segment_count(audio, "b200") == 6
    got = store.load_segment_embeddings(audio, "b200")
    assert got.emb.dtype == np.float16 and got.labels == sorted({long_label, "S1", "S2"})
    order = np.argsort([got.labels.index(l) for l in labels], kind="stable")
    assert np.array_equal(got.emb, emb.astype(np.float16)[order])
    store.save_segment_embeddings(audio, "b200", emb, labels)
    assert store.load_segment_embeddings(audio, "b200").emb.dtype == np.float32
    with pytest.raises(ValueError):
        store.save_segment_embeddings(audio, "b200", emb, labels, dtype=np.int8)
