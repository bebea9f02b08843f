"""k_poolacc2 (cta_group::2: a cluster of two CTAs issues 256-row MMAs and shares every streamed slab of the interleaved
segment matrix) against the oracle, forced through option "cta_group" = 2 on every K-chunk configuration, the column-split
plans, the dense (affinity) mode and the pool-first centroid GEMM; and the single-CTA kernel forced with "cta_group" = 1."""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native, synth
from test_gpu_parity import assert_same, ragged_case, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pairs(ctx):
    ctx.set_option("cta_group", 2)
    yield ctx
    ctx.set_option("cta_group", 0)


@pytest.mark.parametrize("D", [64, 128, 192, 256, 320, 448, 512])
def test_cta2_accumulate_pooling_dims(pairs, oracle, D):
    case = ragged_case(6000 + D, D, P_speakers=500)                    # ~1000 bank rows: several 512-row units, the last ragged
    gpu = run_gpu(pairs, case, 1, 0, 0.354, 10, path=2, acc=2)
    assert pairs.last_path() == (3, 0)
    assert_same(gpu, run_oracle(oracle, case, 1, 0, 0.354, 10), f"cta2 D={D}")


@pytest.mark.parametrize("dtype,thr,k", [(1, -1.0, 10), (0, 0.354, 5), (1, 0.05, 32)])
def test_cta2_modes_and_many_groups(pairs, oracle, dtype, thr, k):
    rng = np.random.default_rng(61)
    counts = rng.integers(0, 40, size=3000)                             # plan A: > 2048 one-column groups, empty ones included
    case = synth.make_case(6100, counts, 700, 192, rows_per_speaker=rng.choice([1, 2, 3], size=700), impostor_frac=0.2)
    gpu = run_gpu(pairs, case, dtype, 0, thr, k, path=2, acc=2)
    assert pairs.last_path()[0] == 3
    assert_same(gpu, run_oracle(oracle, case, dtype, 0, thr, k), "cta2 many groups")


def test_cta2_affinity_dense_mode(pairs, oracle):
    rng = np.random.default_rng(62)
    case = synth.make_case(6200, synth.zipf_counts(rng, 3000, 16), 16, 256, impostor_frac=0.0)
    pairs.set_option("path", 2)
    pairs.set_option("acc", 2)
    try:
        nl, ll = pairs.affinity_pooled(case.seg, case.seg_label, case.G, dtype=1, pool=0)
        assert pairs.last_path()[0] == 3
    finally:
        pairs.set_option("path", 0)
        pairs.set_option("acc", 1)
    ref = oracle.affinity(case.seg, case.goff, mode=1, pool=0)
    np.testing.assert_allclose(nl, ref, rtol=0, atol=2e-5)


def test_auto_takes_pairs_for_many_groups_and_single_cta_can_be_forced(ctx, oracle):
    rng = np.random.default_rng(63)
    counts = rng.integers(1, 6, size=17000)                             # 67 blocks of 256 groups: the auto rule picks CTA pairs
    case = synth.make_case(6300, counts, 600, 64, rows_per_speaker=rng.choice([1, 2], size=600), impostor_frac=0.2)
    ref = run_oracle(oracle, case, 1, 0, 0.354, 4)
    for cg in (0, 1, 2):
        ctx.set_option("cta_group", cg)
        try:
            gpu = run_gpu(ctx, case, 1, 0, 0.354, 4, path=2, acc=2)
        finally:
            ctx.set_option("cta_group", 0)
        assert ctx.last_path() == (3, 0)
        assert_same(gpu, ref, f"cta_group={cg}")


@pytest.mark.parametrize("sched", ["0", "1"])
def test_cta2_group_schedule_covers_every_unit(pairs, oracle, monkeypatch, sched):
    """The static GROUP schedule of the CTA pairs (PaSched: full groups of RB pairs + a last group that deals its blocks round
    robin) must visit every (block, row block) unit exactly once: 20 000 one-column groups = 79 blocks x 3 row blocks on 74
    pairs -> 24 full groups and a last group of 2 pairs, one full super-round of 74 blocks and a partial one."""
    monkeypatch.setenv("SDK_PA_SCHED", sched)
    rng = np.random.default_rng(64)
    counts = rng.integers(1, 4, size=20000)
    case = synth.make_case(6400, counts, 700, 64, rows_per_speaker=rng.choice([1, 2, 3], size=700), impostor_frac=0.2)
    assert 1024 < case.bank.shape[0] <= 1536                            # 3 row blocks of 512 rows
    gpu = run_gpu(pairs, case, 1, 0, 0.354, 4, path=2, acc=2)
    assert pairs.last_path() == (3, 0)
    assert_same(gpu, run_oracle(oracle, case, 1, 0, 0.354, 4), f"group schedule {sched}")
