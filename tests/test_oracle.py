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
