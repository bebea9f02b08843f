"""CPU tests of the oracle itself: pinned against the golden vectors produced by the reference
(tests/golden/*.json, generator tests/golden/make_golden.py) and cross-checked C <-> NumPy."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import matching_np as mnp
from speaker_diarization_toolkit_b200 import synth

GOLD = Path(__file__).parent / "golden"
TRUST_CODE = {"high": 0, "medium": 1, "low": 2, "invalidated": 3, "unknown": 4}


def load(name):
    return json.loads((GOLD / name).read_text())


def test_python_combine_signals_matches_reference_golden():
    """oracle/matching_np.combine_signals == the reference's combine_signals on 409 cases (7 survey KATs + random)."""
    gold = load("combine_signals_golden.json")
    assert len(gold) >= 400
    for case in gold:
        sigs = [mnp.Signal(t, i, s, dict(ev)) for t, i, s, ev in case["signals"]]
        got = mnp.combine_signals("S1", sigs, threshold=case["threshold"])
        exp = case["expect"]
        assert got["speaker_id"] == exp["speaker_id"]
        assert got["confidence"] == exp["confidence"]
        assert got["score"] == exp["score"]            # float64, bit-exact
        assert got["candidates"] == exp["candidates"]
        assert got["signals"] == exp["signals"]


def test_survey_kats_are_in_the_golden_file():
    gold = load("combine_signals_golden.json")
    e = gold[0]["expect"]
    assert (e["speaker_id"], e["confidence"], e["score"]) == ("bob", "low", 0.38)
    assert e["candidates"] == [{"speaker_id": "alice", "score": 0.36000000000000004}]
    e = gold[4]["expect"]          # tie: first inserted wins
    assert e["speaker_id"] == "zed" and e["score"] == 0.32000000000000006
    e = gold[6]["expect"]          # 0.18 >= thr 0.1 but below the "low" band
    assert (e["speaker_id"], e["confidence"]) == ("bob", "unassigned")


def test_c_assign_matches_reference_golden(oracle):
    """orc_assign (min-trust filter + combine_signals over embedding-only rows) == the reference on 400 cases."""
    gold = load("assign_embedding_only_golden.json")
    assert len(gold) == 400
    nonempty = 0
    for case in gold:
        rows = case["rows"]
        k = max(1, len(rows))
        ids = [r["speaker_id"] for r in rows]
        row = np.full((1, k), -1, np.int64)
        score = np.zeros((1, k), np.float32)
        trust = np.full((1, k), 4, np.uint8)
        for i, r in enumerate(rows):
            row[0, i], score[0, i], trust[0, i] = i, r["score"], TRUST_CODE[r["trust_level"]]
        a_idx, a_score, a_conf, c_idx, c_score = oracle.assign(row, score, trust, np.asarray([len(rows)], np.int32),
                                                               case["threshold"], TRUST_CODE[case["min_trust"]])
        exp = case["expect"]
        got_id = None if a_idx[0] < 0 else ids[a_idx[0]]
        assert got_id == exp["speaker_id"]
        assert a_score[0] == exp["score"]                       # float64 bit-exact
        assert ["unassigned", "low", "medium", "high"][a_conf[0]] == exp["confidence"]
        assert [(ids[j], c_score[0, n]) for n, j in enumerate(c_idx[0]) if j >= 0] == \
               [(c["speaker_id"], c["score"]) for c in exp["candidates"]]
        nonempty += exp["speaker_id"] is not None
    assert nonempty >= 50


def test_min_trust_filter_matches_reference_golden():
    for case in load("embedding_signals_golden.json"):
        kept = [row["speaker_id"] for row in case["canned"]
                if row.get("speaker_id") and mnp.passes_min_trust(row.get("trust_level", "unknown"), case["min_trust"])]
        assert kept == [e["speaker_id"] for e in case["expect"]]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("pool", [0, 1])
def test_canonical_c_agrees_with_numpy(oracle, mode, pool):
    """ids identical, |score difference| <= 1e-5 relative (north_star tolerance for fp32)."""
    case = synth.config2(total=600, P=120)
    r, s, c = oracle.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=mode, pool=pool,
                              threshold=0.2, k=5)
    r2, s2, c2 = mnp.identify(case.seg, case.goff, case.bank, case.row_speaker, mode=mode, pool=pool, threshold=0.2, k=5)
    assert np.array_equal(r, r2) and np.array_equal(c, c2)
    np.testing.assert_allclose(s, s2, rtol=1e-5, atol=1e-7)
    hit = case.truth >= 0
    assert np.array_equal(case.row_speaker[r[hit, 0]], case.truth[hit])        # planted speakers are found
    assert (s[~hit] < 0.354).all()                                             # impostors stay below the identify threshold


def test_canonical_normalize_properties(oracle):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 192)).astype(np.float32) * rng.uniform(0.5, 20, (50, 1)).astype(np.float32)
    x[3] = 0
    y, n = oracle.normalize(x, 0)
    assert np.all(y[3] == 0) and n[3] == 0
    np.testing.assert_allclose(np.linalg.norm(y[np.arange(50) != 3], axis=1), 1.0, rtol=2e-7)
    np.testing.assert_allclose(y, mnp.l2_normalize(x), rtol=3e-7, atol=1e-9)
    yb, _ = oracle.normalize(x, 1)
    assert np.array_equal(yb, mnp.bf16_round(y))                               # bf16 RNE identical in C and NumPy
    # scale invariance of everything downstream of the normalise step (power-of-two scale is exact)
    y2, _ = oracle.normalize(x * np.float32(4.0), 0)
    assert np.array_equal(y, y2)


def test_select_semantics(oracle):
    # speaker 0 has rows 0,1; speaker 1 row 2; speaker 2 rows 3,4 (tie inside the speaker -> lowest row)
    sim = np.asarray([[0.5, 0.9, 0.9, 0.7, 0.7], [0.1, 0.2, 0.3, 0.36, 0.1]], np.float32)
    spk = np.asarray([0, 0, 1, 2, 2], np.int32)
    rows, scores, cnt = oracle.select(sim, [3, 3], spk, 3, 0.354, 3)
    assert rows[0].tolist() == [1, 2, 3] and cnt[0] == 3          # 0.9 tie: row 1 before row 2
    assert rows[1].tolist() == [3, -1, -1] and cnt[1] == 1        # only 0.36 >= 0.354
    rows, scores, cnt = oracle.select(sim, [3, 0], spk, 3, -1.0, 2, row_offset=1000)
    assert rows[0].tolist() == [1001, 1002] and cnt.tolist() == [2, 0]   # k cap, global offset, empty label
    r2, s2, c2 = mnp.select_topk(sim, [3, 0], spk, -1.0, 2, row_offset=1000)
    assert np.array_equal(rows, r2) and np.array_equal(cnt, c2)


def test_embedding_only_bounds():
    """SURVEY 8c consequences: an embedding-only score tops out at 0.4 (never above 'medium'); with the default
    assign threshold 0.3 it needs cosine >= 0.75 at high trust and is impossible at medium/low trust."""
    mk = lambda s, t: [mnp.Signal("embedding_match", "a", s, {"trust_level": t})]
    assert mnp.combine_signals("S1", mk(1.0, "high"), 0.3)["confidence"] == "medium"
    assert mnp.combine_signals("S1", mk(0.75, "high"), 0.3)["speaker_id"] == "a"
    assert mnp.combine_signals("S1", mk(0.74, "high"), 0.3)["speaker_id"] is None
    assert mnp.combine_signals("S1", mk(1.0, "medium"), 0.3)["speaker_id"] is None
    assert mnp.combine_signals("S1", mk(1.0, "low"), 0.3)["speaker_id"] is None


# ---- independent float64 restatement of SURVEY 8c steps 1-5 (the tightest pin an unpinned path can get) ---------
def _f64_pooled(seg, goff, bank, mode, pool):
    """Plain NumPy: operands by the spec of step 1-2 (fp64 sum of squares -> fp32 norm -> ONE fp32 reciprocal per row ->
    fp32 multiply -> bf16 RNE in mode 1; the operands are fp32 / bf16 VALUES in every implementation, and a 1-ulp
    different normalisation would flip bf16 roundings), then a float64 GEMM and float64 mean / max per label.
    Shares no code with the C oracle."""
    def ops(x):
        x64 = x.astype(np.float64)
        nrm = np.sqrt((x64 * x64).sum(axis=1)).astype(np.float32)
        inv = np.float32(1.0) / np.maximum(nrm, np.float32(1e-12))
        y = x.astype(np.float32) * inv[:, None]
        return (mnp.bf16_round(y) if mode == 1 else y).astype(np.float64)
    S = ops(seg) @ ops(bank).T                                    # [N, P] float64
    G = len(goff) - 1
    out = np.zeros((G, bank.shape[0]))
    for g in range(G):
        a, b = int(goff[g]), int(goff[g + 1])
        if b > a:
            out[g] = S[a:b].mean(axis=0) if pool == 0 else S[a:b].max(axis=0)
    return out


FIVE_SHAPES = [
    # name, case factory (sliced to sizes the fp64 product finishes in about a second), mode, threshold, k
    ("cfg1", lambda: synth.config1(), 0, 0.354, 3),
    ("cfg2", lambda: synth.config2(total=2000, P=500), 0, 0.354, 10),
    ("cfg3", lambda: synth.config3(recordings=3, seg_per_rec=2000, P=1500, D=192), 1, 0.354, 4),
    ("cfg4", lambda: synth.config4(P=3000, D=512, total=2000, neighbours=11), 1, -1.0, 10),
]


@pytest.mark.parametrize("pool", [0, 1])
@pytest.mark.parametrize("shape", FIVE_SHAPES, ids=[s[0] for s in FIVE_SHAPES])
def test_canonical_agrees_with_independent_float64(oracle, shape, pool):
    """|canonical pooled similarity - float64 restatement| <= 1e-6 on every (label, bank row) cell of the BASELINE config
    shapes (sliced), and the ordered top-k ids agree wherever the float64 scores are separated by more than that."""
    name, make, mode, thr, k = shape
    case = make()
    ref = _f64_pooled(case.seg, case.goff, case.bank, mode, pool)
    seg_ops, _ = oracle.normalize(case.seg, mode)
    bank_ops, _ = oracle.normalize(case.bank, mode)
    got = oracle.pooled(seg_ops, case.goff, bank_ops, pool)
    err = np.abs(got.astype(np.float64) - ref).max()
    assert err <= 1e-6, (name, err)
    rows, scores, cnt = oracle.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=mode, pool=pool,
                                        threshold=thr, k=k)
    # float64 top-k per label over speakers (max over a speaker's rows)
    for g in range(case.G):
        if case.goff[g + 1] == case.goff[g]:
            assert cnt[g] == 0
            continue
        best = np.full(case.n_speakers, -np.inf)
        np.maximum.at(best, case.row_speaker, ref[g])
        order = np.argsort(-best, kind="stable")
        keep = [s for s in order if best[s] >= thr][:k]
        got_spk = case.row_speaker[rows[g, :cnt[g]]].tolist()
        # compare rank by rank until the first float64 near-tie / near-threshold score (below 4e-6 the canonical order rules)
        for i, s in enumerate(keep):
            gap_next = best[s] - best[order[list(order).index(s) + 1]] if list(order).index(s) + 1 < len(order) else 1.0
            if gap_next < 4e-6 or abs(best[s] - thr) < 4e-6:
                break
            assert i < len(got_spk) and got_spk[i] == s, (name, g, i)
            assert abs(float(scores[g, i]) - best[s]) <= 1e-6


def test_canonical_affinity_agrees_with_independent_float64(oracle):
    """config 5 shape (sliced): pooled self-affinity [N, L] against the float64 restatement."""
    case = synth.config5(N=1500, L=16, D=256)
    for pool in (0, 1):
        ref = _f64_pooled(case.seg, case.goff, case.seg, 1, pool).T            # [N, L]
        got = oracle.affinity(case.seg, case.goff, mode=1, pool=pool)
        assert np.abs(got.astype(np.float64) - ref).max() <= 1e-6


def test_threaded_oracle_equals_single_thread(oracle, monkeypatch):
    case = synth.config2(total=1200, P=700)
    seg_ops, _ = oracle.normalize(case.seg, 1)
    bank_ops, _ = oracle.normalize(case.bank, 1)
    monkeypatch.setenv("ORC_THREADS", "1")
    a = oracle.pooled(seg_ops, case.goff, bank_ops, 0)
    monkeypatch.setenv("ORC_THREADS", "5")
    b = oracle.pooled(seg_ops, case.goff, bank_ops, 0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # and equals the one-pair-at-a-time definition
    q = sum(int(oracle.lib().orc_pair_q30(seg_ops[s].ctypes.data_as(__import__("ctypes").c_void_p),
                                          bank_ops[17].ctypes.data_as(__import__("ctypes").c_void_p), seg_ops.shape[1]))
            for s in range(int(case.goff[0]), int(case.goff[1])))
    n = int(case.goff[1] - case.goff[0])
    assert a[0, 17] == np.float32(q / (n * 2.0 ** 30))
