"""Pool-first stage A (option "poolfirst", mean pooling): stage A contracts the label CENTROIDS against the bank -- a
different algorithm from the full segment x bank contraction -- and the certified top-k still returns the oracle's
answer bit for bit, because stage B re-scores the candidates over all the label's segments in the canonical arithmetic."""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native, synth
from test_gpu_parity import assert_same, ragged_case, run_oracle

pytestmark = pytest.mark.gpu


def run_pf(ctx, case, dtype, thr, k, seg=None):
    for key, v in (("path", 2), ("acc", 1), ("gemv", 1), ("cand", 16), ("eps", -1.0), ("poolfirst", 1)):
        ctx.set_option(key, v)
    try:
        ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=dtype)
        out = ctx.identify(case.seg if seg is None else seg, case.seg_label, case.G, pool=0, threshold=thr, k=k)
        path, nfb = ctx.last_path()
    finally:
        ctx.set_option("poolfirst", 0)
        ctx.set_option("path", 0)
    return out, path, nfb


@pytest.mark.parametrize("D", [64, 192, 256, 512])
@pytest.mark.parametrize("dtype,thr,k", [(1, 0.354, 10), (1, -1.0, 10), (0, 0.354, 5)])
def test_pool_first_matches_oracle(ctx, oracle, D, dtype, thr, k):
    case = ragged_case(5000 + D, D, P_speakers=300)
    gpu, path, nfb = run_pf(ctx, case, dtype, thr, k)
    assert path == 5
    if thr > 0.2:
        assert nfb == 0, "thresholded planted data must certify without the exhaustive pass"
    assert_same(gpu, run_oracle(oracle, case, dtype, 0, thr, k), f"pool-first D={D}")


def test_pool_first_config3_slice_and_fp16(ctx, oracle):
    case = synth.config3(recordings=6, seg_per_rec=1500, P=2500, D=192)
    gpu, path, nfb = run_pf(ctx, case, 1, 0.354, 4)
    assert (path, nfb) == (5, 0)
    assert_same(gpu, run_oracle(oracle, case, 1, 0, 0.354, 4), "pool-first cfg3 slice")
    h = case.seg.astype(np.float16)
    gpu16, path, nfb = run_pf(ctx, case, 1, 0.354, 4, seg=h)
    ref16 = oracle.identify(h.astype(np.float32), case.goff, case.bank, case.row_speaker, case.n_speakers, mode=1, pool=0, threshold=0.354, k=4)
    assert path == 5
    assert_same(gpu16, ref16, "pool-first fp16 input")


def test_pool_first_giant_all_positive_label(ctx, oracle):
    """one label of 30 000 all-positive segments (every partial sum grows): the centroid's fp32 accumulation error must stay
    inside the margin model -- measured here against the canonical pooled scores of the candidates"""
    rng = np.random.default_rng(31)
    D, n = 128, 30000
    cent = np.abs(rng.standard_normal(D))
    cent /= np.linalg.norm(cent)
    seg = np.abs(cent[None, :] + 0.35 * rng.standard_normal((n + 50, D)) / np.sqrt(D)).astype(np.float32)
    bank = np.abs(cent[None, :] + rng.uniform(0.05, 1.5, (4000, 1)) * rng.standard_normal((4000, D)) / np.sqrt(D)).astype(np.float32)
    goff = np.asarray([0, n, n + 50], np.int64)
    lab = np.repeat(np.arange(2, dtype=np.int32), [n, 50])
    spk = np.arange(len(bank), dtype=np.int32)
    for key, v in (("path", 2), ("cand", 16), ("eps", -1.0), ("poolfirst", 1)):
        ctx.set_option(key, v)
    try:
        ctx.bank_load(bank, spk, None, dtype=1)
        rows, scores, counts = ctx.identify(seg, lab, 2, pool=0, threshold=-1.0, k=10)
        assert ctx.last_path()[0] == 5
        ref = oracle.identify(seg, goff, bank, spk, len(bank), mode=1, pool=0, threshold=-1.0, k=10)
        assert np.array_equal(rows, ref[0]) and np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32))
        cand, approx, eps_base, eps_chain = ctx.stage_a()
        seg_ops, _ = oracle.normalize(seg, 1)
        bank_ops, _ = oracle.normalize(bank, 1)
        r = cand[0][cand[0] >= 0]
        can = oracle.pooled(seg_ops[:n], np.asarray([0, n], np.int64), bank_ops[r], 0)[0]
        worst = float(np.abs(approx[0][cand[0] >= 0] - can).max())
        eps_g = eps_base + eps_chain * (n + 70)
        print(f"\\n[certificate] pool-first, one label of {n} all-positive segments: measured {worst:.3e}, model {eps_g:.3e}")
        assert worst <= 0.6 * eps_g, (worst, eps_g)
    finally:
        ctx.set_option("poolfirst", 0)
        ctx.set_option("path", 0)
