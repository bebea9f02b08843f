"""Randomised parity: seeded random shapes pushed through every stage-A kernel and the exact path, bit-exact against the
C oracle (ids, order, counts AND scores).  Complements the hand-picked cases of test_gpu_parity.py."""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import synth

pytestmark = pytest.mark.gpu

DIMS = [64, 100, 128, 192, 256, 320, 384, 512]
SIZES = [0, 1, 2, 3, 15, 16, 17, 31, 33, 64, 100, 255, 256, 257, 600, 1500]


def random_case(seed):
    rng = np.random.default_rng(seed)
    D = int(rng.choice(DIMS))
    G = int(rng.integers(1, 41))
    counts = rng.choice(SIZES, size=G, p=None)
    if rng.random() < 0.25:                     # a handful of segments only: the bank-stream kernel's regime
        counts = np.zeros(G, dtype=np.int64)
        for _ in range(int(rng.integers(1, 9))):
            counts[rng.integers(0, G)] += 1
    if counts.sum() == 0:
        counts[0] = 5
    n_spk = int(rng.integers(5, 500))
    rps = rng.choice([1, 2, 3], size=n_spk)
    case = synth.make_case(seed, counts, n_spk, D, rows_per_speaker=rps, impostor_frac=float(rng.choice([0.0, 0.2, 0.5])),
                           neighbours=int(rng.choice([0, 3])))
    dtype = int(rng.integers(0, 2))
    pool = int(rng.integers(0, 2))
    thr = float(rng.choice([-1.0, 0.2, 0.354, 0.6]))
    k = int(rng.choice([1, 4, 10, 32]))
    return case, dtype, pool, thr, k


@pytest.mark.parametrize("seed", range(40))
def test_random_shapes_all_paths(ctx, oracle, seed):
    case, dtype, pool, thr, k = random_case(7000 + seed)
    ref = oracle.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=dtype, pool=pool, threshold=thr, k=k)
    # (path, acc, gemv, cta_group, poolfirst): auto, exact, generic tcgen05, accumulate-pooling (forced; mean pooling only)
    # on single CTAs and on CTA pairs, bank stream, pool-first (mean pooling only; also on CTA pairs)
    variants = [(0, 1, 1, 0, 0), (1, 0, 0, 0, 0), (2, 0, 0, 0, 0), (2, 2, 0, 1, 0), (2, 2, 0, 2, 0), (2, 1, 1, 0, 0), (2, 1, 0, 0, 1),
                (2, 1, 0, 2, 1)]
    N = int(case.goff[-1])
    for path, acc, gemv, cta_group, poolfirst in variants:
        ctx.set_option("path", path)
        ctx.set_option("acc", acc)
        ctx.set_option("gemv", gemv)
        ctx.set_option("cta_group", cta_group)
        ctx.set_option("poolfirst", poolfirst)
        ctx.set_option("cand", 16)
        ctx.set_option("eps", -1.0)
        ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=dtype)
        rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, pool=pool, threshold=thr, k=k)
        what = f"seed {seed} D={case.seg.shape[1]} G={case.G} N={N} P={case.bank.shape[0]} dtype={dtype} pool={pool} thr={thr} k={k} " \
               f"variant={(path, acc, gemv, cta_group, poolfirst)} took path {ctx.last_path()}"
        assert np.array_equal(counts, ref[2]), what
        assert np.array_equal(rows, ref[0]), what
        assert np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32)), what
        took = ctx.last_path()[0]
        if path == 1:
            assert took == 1
        if path == 2 and acc == 2 and pool == 0 and gemv == 0 and case.seg.shape[1] % 4 == 0:
            assert took == 3, what
        if path == 2 and gemv == 1 and N <= 8 and case.G <= 64:
            assert took == 4, what
        if poolfirst and pool == 0 and case.seg.shape[1] % 4 == 0:
            assert took == 5, what
    ctx.set_option("path", 0)
    ctx.set_option("acc", 1)
    ctx.set_option("gemv", 1)
    ctx.set_option("cta_group", 0)
    ctx.set_option("poolfirst", 0)
