"""The small-query latency path (csrc/gemv.cu: <= 8 query segments, two kernels, no host round trip inside the call):
shapes the other suites do not reach -- tiny banks, several selection passes per CTA, many labels on one list slot --
and the deferred settling of label errors and failed certificates."""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native, synth
from test_gpu_parity import assert_same, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


def small_case(seed, counts, P_speakers, D, rps=(1, 2, 3), **kw):
    rng = np.random.default_rng(seed)
    return synth.make_case(seed, counts, P_speakers, D, rows_per_speaker=rng.choice(rps, size=P_speakers), impostor_frac=0.25, **kw)


@pytest.mark.parametrize("P_speakers,D,counts,dtype,pool,thr,k", [
    (5, 64, [3, 2], 1, 0, -1.0, 10),                      # a bank smaller than one 32-row tile
    (17, 128, [1, 0, 7], 1, 1, 0.2, 3),                   # 33..50 rows: two tiles, the second ragged
    (40000, 64, [2, 1, 1, 1, 1, 1, 1], 1, 0, -1.0, 10),   # ~80k rows: every CTA streams ~17 tiles
    (110000, 64, [4, 4], 1, 0, 0.354, 10),                # ~220k rows: more than 32 tiles per CTA -> two selection passes
    (110000, 64, [8], 0, 1, -1.0, 32),                    # fp32 bank, one label of 8, max pooling, k = 32
    (3000, 512, [1] * 8, 1, 0, -1.0, 10),
])
def test_small_path_shapes(ctx, oracle, P_speakers, D, counts, dtype, pool, thr, k):
    case = small_case(4000 + P_speakers + D, counts, P_speakers, D, neighbours=3)
    gpu = run_gpu(ctx, case, dtype, pool, thr, k, path=2, gemv=1)
    path, nfb = ctx.last_path()
    assert path == 4
    assert_same(gpu, run_oracle(oracle, case, dtype, pool, thr, k), f"small P={case.bank.shape[0]} D={D}")


def test_small_path_many_labels_sparse(ctx, oracle):
    """64 label groups, only a few of them hold segments: results of the empty ones are empty, the others exact"""
    counts = [0] * 64
    for g, n in ((3, 2), (17, 1), (40, 3), (63, 2)):
        counts[g] = n
    case = small_case(4100, counts, 900, 192)
    gpu = run_gpu(ctx, case, 1, 0, 0.354, 5, path=2, gemv=1)
    assert ctx.last_path()[0] == 4
    assert_same(gpu, run_oracle(oracle, case, 1, 0, 0.354, 5), "sparse labels")


def test_small_path_failed_certificate_is_settled_at_fetch(ctx, oracle):
    """A hopeless margin makes every certificate fail: the call itself does not notice (no host round trip), the fetch
    settles the labels exhaustively -- and redoes the assignment that was computed from the unsettled lists."""
    case = small_case(4200, [2, 3, 1], 600, 128, neighbours=4)
    ctx.set_option("path", 2)
    ctx.set_option("gemv", 1)
    ctx.set_option("cand", 1)
    ctx.set_option("eps", 0.5)
    try:
        ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=1)
        rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, pool=0, threshold=-1.0, k=10)
        path, nfb = ctx.last_path()
        assert path == 4 and nfb == 3
        ref = run_oracle(oracle, case, 1, 0, -1.0, 10)
        assert_same((rows, scores, counts), ref, "settled")
        # device pointers: identify_dev + assign return at once; fetch settles and re-assigns
        torch = pytest.importorskip("torch")
        seg = torch.from_numpy(case.seg).cuda()
        lab = torch.from_numpy(case.seg_label.astype(np.int32)).cuda()
        torch.cuda.synchronize()
        ctx.identify_dev(seg.data_ptr(), lab.data_ptr(), seg.shape[0], case.G, 0, -1.0, 10)
        ctx.assign(0.2, "low")
        out = ctx.fetch(with_assign=True)
        assert_same((out["row"], out["score"], out["count"]), ref, "settled (dev)")
        a = oracle.assign(out["row"], out["score"], out["trust"], out["count"], 0.2, 2)
        assert np.array_equal(out["assign_idx"], a[0]) and np.array_equal(out["assign_score"], a[1])
        assert ctx.last_path() == (4, 3)
    finally:
        ctx.set_option("cand", 16)
        ctx.set_option("eps", -1.0)
        ctx.set_option("path", 0)


def test_small_path_label_errors_surface_at_fetch(ctx):
    case = small_case(4300, [2, 2], 300, 64)
    ctx.set_option("path", 2)
    ctx.set_option("gemv", 1)
    try:
        ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=1)
        with pytest.raises(_native.NativeError, match="non-decreasing"):
            ctx.identify(case.seg, case.seg_label[::-1].copy(), case.G)
        with pytest.raises(_native.NativeError, match="out of range"):
            ctx.identify(case.seg, case.seg_label + 7, case.G)
        torch = pytest.importorskip("torch")
        seg = torch.from_numpy(case.seg).cuda()
        lab = torch.from_numpy((case.seg_label + 7).astype(np.int32)).cuda()
        torch.cuda.synchronize()
        ctx.identify_dev(seg.data_ptr(), lab.data_ptr(), seg.shape[0], case.G, 0, 0.354, 4)      # queued, not checked yet
        with pytest.raises(_native.NativeError, match="out of range"):
            ctx.fetch()
        # and the context is usable afterwards
        rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, k=4)
        assert ctx.last_path()[0] == 4 and counts.shape == (2,)
    finally:
        ctx.set_option("path", 0)


def test_small_path_fp16_queries(ctx, oracle):
    case = small_case(4400, [3, 1, 2], 2000, 256)
    h = case.seg.astype(np.float16)
    ctx.set_option("path", 2)
    ctx.set_option("gemv", 1)
    try:
        ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=1)
        gpu = ctx.identify(h, case.seg_label, case.G, pool=0, threshold=0.354, k=6)
        assert ctx.last_path()[0] == 4
        ref = oracle.identify(h.astype(np.float32), case.goff, case.bank, case.row_speaker, case.n_speakers, mode=1, pool=0, threshold=0.354, k=6)
        assert_same(gpu, ref, "small f16")
    finally:
        ctx.set_option("path", 0)
