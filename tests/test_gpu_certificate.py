"""Adversarial tests of the certified top-k.

The tensor-core stage A only PROPOSES candidates; ids are bit-exact because (a) the candidates are re-scored in the
canonical arithmetic and (b) a certificate proves that no row left out can enter the result: its stage-A score plus
eps_g must stay below the k-th kept score (or the threshold).  eps_g is a MODEL of the stage-A rounding error (api.cu /
DESIGN.md section 2: 6 * 2^-23 per accumulator update relative to the running magnitude).  These tests attack exactly that:
all-positive embeddings (every partial sum grows monotonically: the worst case for a truncating accumulator), label
groups of thousands of segments accumulated in ONE TMEM column, and dozens of bank rows whose canonical pooled scores
differ by less than 1e-6 around the k-th place.  They assert the ids / order / scores against the oracle AND that the
measured |stage A - canonical| stays below the model with a margin.
"""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native

pytestmark = pytest.mark.gpu


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def _near_tie_bank(oracle, rng, seg_giant, D, n_pool=24000, n_tie=40, n_above=5, n_below=600):
    """Bank whose canonical pooled scores against `seg_giant` hold `n_above` clear winners, then `n_tie` rows inside a
    window far narrower than any stage-A error, then `n_below` clear losers.  The window is found, not constructed:
    thousands of 1-ulp(bf16) perturbations of one row, canonical scores from the oracle, densest run of n_tie."""
    cent = _unit(np.abs(seg_giant).mean(axis=0))
    base = _unit(np.abs(cent + 0.3 * rng.standard_normal(D) / np.sqrt(D)))      # room above for the clear winners
    pert = np.ones((n_pool, D), np.float32)
    for i in range(n_pool):
        idx = rng.choice(D, size=16, replace=False)
        pert[i, idx] += rng.choice([-1.0, 1.0], size=16).astype(np.float32) * np.float32(2.0 ** -8)
    pool_rows = (base[None, :].astype(np.float32) * pert).astype(np.float32)
    ops, _ = oracle.normalize(pool_rows, 1)
    seg_ops, _ = oracle.normalize(seg_giant, 1)
    goff = np.asarray([0, len(seg_giant)], np.int64)
    sc = oracle.pooled(seg_ops, goff, ops, 0)[0].astype(np.float64)
    order = np.argsort(sc, kind="stable")
    span = sc[order[n_tie - 1:]] - sc[order[:len(order) - n_tie + 1]]
    j = int(np.argmin(span))
    tie_idx = order[j:j + n_tie]
    tie_span = float(span[j])
    level = float(sc[tie_idx].mean())
    # clear winners: closer to the centroid; clear losers: noisier rows (all positive)
    above = _unit(np.abs(cent[None, :] + 0.01 * rng.standard_normal((64, D)) / np.sqrt(D))).astype(np.float32)
    below = _unit(np.abs(cent[None, :] + rng.uniform(0.3, 2.0, (n_below * 2, 1)) * rng.standard_normal((n_below * 2, D)) / np.sqrt(D))).astype(np.float32)
    def scores(rows):
        o, _ = oracle.normalize(rows, 1)
        return oracle.pooled(seg_ops, goff, o, 0)[0].astype(np.float64)
    sa, sb = scores(above), scores(below)
    above = above[sa > level + 2e-3][:n_above]
    below = below[sb < level - 2e-3][:n_below]
    assert len(above) == n_above and len(below) >= n_below // 2
    bank = np.concatenate([above, pool_rows[tie_idx], below]).astype(np.float32)
    bank = bank[rng.permutation(len(bank))]
    bank *= rng.uniform(0.5, 20.0, (len(bank), 1)).astype(np.float32)          # the normalise kernel does real work
    return bank, tie_span


def _check(ctx, oracle, seg, goff, bank, pool, k, what):
    lab = np.repeat(np.arange(len(goff) - 1, dtype=np.int32), np.diff(goff))
    spk = np.arange(len(bank), dtype=np.int32)
    ctx.bank_load(bank, spk, None, dtype=_native.DTYPE_BF16)
    rows, scores, counts = ctx.identify(seg, lab, len(goff) - 1, pool=pool, threshold=-1.0, k=k)
    ref = oracle.identify(seg, goff, bank, spk, len(bank), mode=1, pool=pool, threshold=-1.0, k=k)
    assert np.array_equal(counts, ref[2]), what
    assert np.array_equal(rows, ref[0]), f"{what}: ids / order differ from the oracle"
    assert np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32)), what
    return rows


def _stage_a_error(ctx, oracle, seg, goff, bank, pool, groups):
    """max |stage-A pooled score - canonical| over the first-chance candidates of `groups`, and the model's eps_g."""
    cand, approx, eps_base, eps_chain = ctx.stage_a()
    seg_ops, _ = oracle.normalize(seg, 1)
    bank_ops, _ = oracle.normalize(bank, 1)
    worst = 0.0
    for g in groups:
        r = cand[g][cand[g] >= 0]
        can = oracle.pooled(seg_ops[goff[g]:goff[g + 1]], np.asarray([0, goff[g + 1] - goff[g]], np.int64), bank_ops[r], pool)[0]
        worst = max(worst, float(np.abs(approx[g][cand[g] >= 0].astype(np.float64) - can.astype(np.float64)).max()))
    return worst, eps_base, eps_chain


def test_certificate_adversarial_accumulate_pooling_long_chain(ctx, oracle):
    """One label of 3 000 all-positive segments accumulated in a single TMEM column (plan A: > 2048 labels, c = 1) against
    a bank with 40 rows inside a sub-1e-6 window at the k-th place."""
    rng = np.random.default_rng(4242)
    D, n_giant, n_small = 64, 3000, 2100
    cent = _unit(np.abs(rng.standard_normal(D)))
    seg_giant = _unit(np.abs(cent[None, :] + 0.35 * rng.standard_normal((n_giant, D)) / np.sqrt(D))).astype(np.float32)
    bank, tie_span = _near_tie_bank(oracle, rng, seg_giant, D)
    assert tie_span < 1e-6, tie_span
    seg_small = _unit(np.abs(cent[None, :] + 0.5 * rng.standard_normal((n_small, D)) / np.sqrt(D))).astype(np.float32)
    seg = np.concatenate([seg_giant, seg_small]) * rng.uniform(0.5, 20.0, (n_giant + n_small, 1)).astype(np.float32)
    goff = np.concatenate([[0], [n_giant], n_giant + 1 + np.arange(n_small)]).astype(np.int64)
    ctx.set_option("path", 2)
    ctx.set_option("acc", 2)
    ctx.set_option("gemv", 0)
    ctx.set_option("cand", 16)
    ctx.set_option("eps", -1.0)
    try:
        _check(ctx, oracle, seg, goff, bank, 0, 10, "accumulate-pooling, giant all-positive label")
        assert ctx.last_path()[0] == 3
        # the near-tie window straddles the k-th place: the first list cannot be certified
        assert ctx.last_retry() + ctx.last_path()[1] >= 1
        worst, eps_base, eps_chain = _stage_a_error(ctx, oracle, seg, goff, bank, 0, [0, 1, 2, 3])
        eps_giant = eps_base + eps_chain * n_giant
        print(f"\n[certificate] accumulate-pooling chain {n_giant} x {D // 16} updates: measured max |stage A - canonical| = {worst:.3e}, "
              f"model eps_g = {eps_giant:.3e} (ratio {eps_giant / max(worst, 1e-12):.1f}), near-tie window {tie_span:.2e}")
        assert worst <= 0.6 * eps_giant, (worst, eps_giant)
    finally:
        ctx.set_option("path", 0)
        ctx.set_option("acc", 1)
        ctx.set_option("gemv", 1)


@pytest.mark.parametrize("pool", [0, 1])
def test_certificate_adversarial_epilogue_pooling(ctx, oracle, pool):
    """Generic tcgen05 kernel (pooling in the epilogue), D = 512, labels of 2 000 all-positive segments, near-tie bank."""
    rng = np.random.default_rng(777 + pool)
    D, n = 512, 2000
    cent = _unit(np.abs(rng.standard_normal(D)))
    segs = [_unit(np.abs(cent[None, :] + s * rng.standard_normal((n, D)) / np.sqrt(D))).astype(np.float32) for s in (0.35, 0.5, 0.8)]
    bank, tie_span = _near_tie_bank(oracle, rng, segs[0], D, n_pool=6000, n_below=900)
    seg = np.concatenate(segs)
    goff = np.asarray([0, n, 2 * n, 3 * n], np.int64)
    ctx.set_option("path", 2)
    ctx.set_option("acc", 0)
    ctx.set_option("gemv", 0)
    ctx.set_option("cand", 16)
    ctx.set_option("eps", -1.0)
    try:
        _check(ctx, oracle, seg, goff, bank, pool, 10, f"epilogue pooling, pool={pool}")
        assert ctx.last_path()[0] == 2
        worst, eps_base, eps_chain = _stage_a_error(ctx, oracle, seg, goff, bank, pool, [0, 1, 2])
        eps_g = eps_base + eps_chain * (n // 32 + 70 if pool == 0 else 0)
        print(f"\n[certificate] epilogue pooling (pool={pool}), D=512, n={n}: measured {worst:.3e}, model {eps_g:.3e} "
              f"(ratio {eps_g / max(worst, 1e-12):.1f}), near-tie window {tie_span:.2e}")
        assert worst <= 0.6 * eps_g, (worst, eps_g)
    finally:
        ctx.set_option("path", 0)
        ctx.set_option("acc", 1)
        ctx.set_option("gemv", 1)


def test_certificate_adversarial_bank_stream(ctx, oracle):
    """Bank-stream kernel (<= 8 query segments, mma.sync): all-positive centroids, near-tie bank, D = 512."""
    rng = np.random.default_rng(99)
    D = 512
    cent = _unit(np.abs(rng.standard_normal(D)))
    seg8 = _unit(np.abs(cent[None, :] + 0.2 * rng.standard_normal((8, D)) / np.sqrt(D))).astype(np.float32)
    bank, tie_span = _near_tie_bank(oracle, rng, seg8[:3], D, n_pool=6000, n_below=3000)
    goff = np.asarray([0, 3, 4, 8], np.int64)
    ctx.set_option("path", 2)
    ctx.set_option("gemv", 1)
    ctx.set_option("cand", 16)
    ctx.set_option("eps", -1.0)
    try:
        _check(ctx, oracle, seg8, goff, bank, 0, 10, "bank stream")
        assert ctx.last_path()[0] == 4
        worst, eps_base, eps_chain = _stage_a_error(ctx, oracle, seg8, goff, bank, 0, [0, 1, 2])
        print(f"\n[certificate] bank stream, D=512: measured {worst:.3e}, model {eps_base:.3e} (ratio {eps_base / max(worst, 1e-12):.1f})")
        assert worst <= 0.6 * eps_base, (worst, eps_base)
    finally:
        ctx.set_option("path", 0)


def test_fp32_bank_stage_a_error_is_inside_the_margin(ctx, oracle):
    """fp32 bank: stage A rounds both operands to bf16; the margin must cover that (2^-8 relative) too."""
    rng = np.random.default_rng(5)
    D, n = 256, 600
    cent = _unit(np.abs(rng.standard_normal(D)))
    seg = _unit(np.abs(cent[None, :] + 0.35 * rng.standard_normal((2 * n, D)) / np.sqrt(D))).astype(np.float32)
    bank = _unit(np.abs(cent[None, :] + rng.uniform(0.05, 1.5, (3000, 1)) * rng.standard_normal((3000, D)) / np.sqrt(D))).astype(np.float32)
    goff = np.asarray([0, n, 2 * n], np.int64)
    lab = np.repeat(np.arange(2, dtype=np.int32), n)
    spk = np.arange(len(bank), dtype=np.int32)
    ctx.set_option("path", 2)
    ctx.set_option("acc", 0)
    ctx.set_option("eps", -1.0)
    try:
        ctx.bank_load(bank, spk, None, dtype=_native.DTYPE_F32)
        rows, scores, counts = ctx.identify(seg, lab, 2, pool=0, threshold=-1.0, k=10)
        ref = oracle.identify(seg, goff, bank, spk, len(bank), mode=0, pool=0, threshold=-1.0, k=10)
        assert np.array_equal(rows, ref[0]) and np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32))
        cand, approx, eps_base, eps_chain = ctx.stage_a()
        seg_ops, _ = oracle.normalize(seg, 0)
        bank_ops, _ = oracle.normalize(bank, 0)
        worst = 0.0
        for g in range(2):
            r = cand[g][cand[g] >= 0]
            can = oracle.pooled(seg_ops[goff[g]:goff[g + 1]], np.asarray([0, n], np.int64), bank_ops[r], 0)[0]
            worst = max(worst, float(np.abs(approx[g][cand[g] >= 0] - can).max()))
        print(f"\n[certificate] fp32 bank, bf16 stage A: measured {worst:.3e}, model {eps_base:.3e}")
        assert worst <= 0.5 * eps_base
    finally:
        ctx.set_option("path", 0)
        ctx.set_option("acc", 1)
