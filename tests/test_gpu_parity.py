"""GPU parity tests: the CUDA path, called THROUGH THE C-ABI, against the C oracle on identical inputs.

Bar (BASELINE.json): matched rows / ids / counts bit-exact; scores here are bit-exact too, because the
CUDA path re-scores in the oracle's canonical arithmetic (oracle/canonical.c).
"""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native, synth

pytestmark = pytest.mark.gpu


def run_gpu(ctx, case, dtype, pool, thr, k, path, cand=None, eps=None, acc=0, gemv=0):
    """acc: 0 = generic tcgen05 kernel (pooling in the epilogue), 1 = auto, 2 = force accumulate-pooling;
    gemv: 1 = bank-stream kernel when there are <= 8 query segments."""
    ctx.set_option("path", path)
    ctx.set_option("acc", acc)
    ctx.set_option("gemv", gemv)
    ctx.set_option("cand", cand if cand is not None else 16)
    ctx.set_option("eps", eps if eps is not None else -1.0)
    ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=dtype)
    rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, pool=pool, threshold=thr, k=k)
    return rows, scores, counts


def run_oracle(oracle, case, dtype, pool, thr, k):
    return oracle.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=dtype, pool=pool,
                           threshold=thr, k=k)


def assert_same(gpu, ref, what=""):
    (r, s, c), (orow, oscore, ocount) = gpu, ref
    assert np.array_equal(c, ocount), f"{what}: counts differ"
    assert np.array_equal(r, orow), f"{what}: matched rows differ\n{r[:4]}\n{orow[:4]}"
    assert np.array_equal(s.view(np.uint32), oscore.view(np.uint32)), f"{what}: scores not bit-identical"


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("pool", [0, 1])
def test_config1_exact(ctx, oracle, dtype, pool):
    case = synth.config1()
    gpu = run_gpu(ctx, case, dtype, pool, 0.354, 3, path=1)
    assert_same(gpu, run_oracle(oracle, case, dtype, pool, 0.354, 3), "cfg1")
    assert ctx.last_path()[0] == 1
    # planted truth: S1 -> speaker 0, S2 -> speaker 1
    assert gpu[0][0, 0] == 0 and gpu[0][1, 0] == 1


@pytest.mark.parametrize("dtype,pool,thr,k", [(0, 0, 0.354, 10), (0, 1, 0.354, 10), (1, 0, 0.354, 10), (0, 0, -1.0, 10),
                                              (1, 1, 0.1, 32), (0, 0, 0.95, 1)])
def test_config2_exact(ctx, oracle, dtype, pool, thr, k):
    case = synth.config2()
    gpu = run_gpu(ctx, case, dtype, pool, thr, k, path=1)
    assert_same(gpu, run_oracle(oracle, case, dtype, pool, thr, k), "cfg2")


def test_auto_path_small_is_exact(ctx):
    case = synth.config2()
    run_gpu(ctx, case, 0, 0, 0.354, 10, path=0)
    assert ctx.last_path()[0] == 1


def ragged_case(seed, D, P_speakers=150, rps=(1, 2, 3)):
    rng = np.random.default_rng(seed)
    counts = [1, 15, 16, 17, 0, 63, 64, 65, 300, 0, 1000, 129, 2, 127]
    rows_per = rng.choice(rps, size=P_speakers)
    return synth.make_case(seed, counts, P_speakers, D, rows_per_speaker=rows_per, impostor_frac=0.2)


@pytest.mark.parametrize("D", [64, 100, 128, 192, 256, 320, 384, 448, 512])
def test_tensor_path_ragged_dims(ctx, oracle, D):
    """tcgen05 path on every K-chunk configuration, ragged groups (empty, 1, straddling 16/64 boundaries),
    bank size not a multiple of 128, several rows per speaker, bf16 operands."""
    case = ragged_case(1000 + D, D)
    gpu = run_gpu(ctx, case, 1, 0, 0.354, 10, path=2)
    assert ctx.last_path() == (2, 0), "tcgen05 path must certify every label (no exhaustive fallback)"
    assert_same(gpu, run_oracle(oracle, case, 1, 0, 0.354, 10), f"tc D={D}")


@pytest.mark.parametrize("dtype,pool,thr,k", [(1, 1, 0.354, 10), (0, 0, 0.354, 10), (0, 1, 0.5, 3), (1, 0, -1.0, 10),
                                              (1, 1, -1.0, 10), (1, 0, 0.05, 32)])
def test_tensor_path_modes(ctx, oracle, dtype, pool, thr, k):
    case = ragged_case(77, 192, P_speakers=400)
    gpu = run_gpu(ctx, case, dtype, pool, thr, k, path=2)
    path, nfb = ctx.last_path()
    assert path == 2
    if thr > 0.2:
        assert nfb == 0, "thresholded planted data must certify without fallback"
    assert_same(gpu, run_oracle(oracle, case, dtype, pool, thr, k), "tc modes")


def test_tensor_path_certificate_fallback(ctx, oracle):
    """A hopeless certificate (one candidate per label, huge eps) must route the groups through the
    exhaustive canonical fallback and still give the oracle's answer."""
    case = ragged_case(78, 128, P_speakers=300)
    gpu = run_gpu(ctx, case, 1, 0, -1.0, 10, path=2, cand=1, eps=0.5)
    assert ctx.last_path()[1] > 0, "expected certificate failures"
    assert_same(gpu, run_oracle(oracle, case, 1, 0, -1.0, 10), "fallback")


@pytest.mark.parametrize("D,dtype,pool,thr,k,counts", [
    (512, 1, 0, -1.0, 10, [1] * 8),                        # config 4 variant (i): 8 pooled label centroids
    (192, 1, 0, 0.354, 4, [2, 0, 1, 3, 0, 0, 1]),
    (256, 1, 1, 0.2, 10, [5, 3]),
    (64, 1, 0, 0.354, 3, [1]),
    (320, 0, 0, 0.354, 5, [0, 4, 4, 0]),                   # fp32 bank: stage A on the bf16 copy, re-score in fp32
    (448, 1, 1, -1.0, 32, [1, 1, 1, 2]),
])
def test_bank_stream_gemv_path(ctx, oracle, D, dtype, pool, thr, k, counts):
    """<= 8 query segments: stage A is one HBM-bound pass over the bank on the CUDA cores (path 4); same certified
    top-k, same oracle, bit-exact."""
    rng = np.random.default_rng(D + pool)
    case = synth.make_case(1200 + D, counts, 700, D, rows_per_speaker=rng.choice([1, 2, 3], size=700), impostor_frac=0.25, neighbours=3)
    gpu = run_gpu(ctx, case, dtype, pool, thr, k, path=2, gemv=1)
    path, nfb = ctx.last_path()
    assert path == 4, "expected the bank-stream kernel"
    if thr > 0.3:
        assert nfb == 0
    assert_same(gpu, run_oracle(oracle, case, dtype, pool, thr, k), f"gemv D={D}")


def test_tensor_path_certificate_second_chance(ctx, oracle):
    """No threshold and few candidates: some certificates fail on the first candidate list; the second, wider list
    (64 candidates from the slots the GEMM left behind) settles them without the exhaustive pass."""
    case = ragged_case(79, 192, P_speakers=900)
    gpu = run_gpu(ctx, case, 1, 0, -1.0, 10, path=2, cand=10, eps=0.02)      # eps far above the real error: forces retries
    assert ctx.last_retry() > 0, "expected first-list certificate failures"
    assert ctx.last_path()[1] < ctx.last_retry()
    assert_same(gpu, run_oracle(oracle, case, 1, 0, -1.0, 10), "second chance")


def test_tensor_path_large_bank_many_rowblocks(ctx, oracle):
    rng = np.random.default_rng(5)
    case = synth.make_case(5, synth.zipf_counts(rng, 600, 6), 2100, 192, neighbours=5, impostor_frac=0.0)
    gpu = run_gpu(ctx, case, 1, 0, 0.354, 10, path=2)
    assert ctx.last_path() == (2, 0)
    assert_same(gpu, run_oracle(oracle, case, 1, 0, 0.354, 10), "tc many row blocks")
    assert (gpu[2] >= 2).any(), "planted neighbours should give multi-entry top-k lists"


def many_groups_case(seed, D, G=300, P_speakers=260, max_size=120):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, max_size, size=G)
    counts[rng.integers(0, G, size=12)] = 0            # empty labels
    counts[:3] = [1, 400, 257]                         # a singleton, and groups longer than one accumulator tile is wide
    rps = rng.choice([1, 2, 3], size=P_speakers)
    return synth.make_case(seed, counts, P_speakers, D, rows_per_speaker=rps, impostor_frac=0.15, neighbours=3)


def test_accumulate_pooling_auto_choice(ctx):
    """auto mode: D <= 256 (the generic kernel is epilogue-bound there) or thousands of groups -> accumulate-pooling
    (path 3), also for a few outsized groups, whose columns are split; D = 512 with few groups, and max pooling,
    stay on the generic tcgen05 kernel (path 2)."""
    even = synth.make_case(1, [40] * 512, 100, 64)
    run_gpu(ctx, even, 1, 0, 0.354, 3, path=2, acc=1)
    assert ctx.last_path()[0] == 3
    skew = synth.make_case(2, [2000] + [10] * 511, 100, 64)
    run_gpu(ctx, skew, 1, 0, 0.354, 3, path=2, acc=1)
    assert ctx.last_path()[0] == 3
    run_gpu(ctx, skew, 1, 1, 0.354, 3, path=2, acc=1)
    assert ctx.last_path()[0] == 2
    wide = synth.make_case(3, [40] * 64, 100, 512)
    run_gpu(ctx, wide, 1, 0, 0.354, 3, path=2, acc=1)
    assert ctx.last_path()[0] == 2


@pytest.mark.parametrize("D,dtype,thr,k,counts", [
    (256, 1, 0.354, 10, [250, 251, 0, 249, 260, 1, 244, 255]),                 # one meeting, 8 labels (config 2 shape)
    (192, 1, 0.354, 4, [5000, 0, 3, 700, 64, 65, 1, 1300, 2, 256, 257, 40000]),  # outsized groups: up to 256 columns each
    (64, 1, -1.0, 10, [33] * 40 + [0, 1, 2, 3]),
    (512, 1, 0.2, 8, [900, 30, 31, 1200, 7]),
    (128, 0, 0.354, 5, [400, 1, 0, 2000, 129, 77]),                            # fp32 bank: re-score from the plain fp32 copy
    (320, 1, 0.354, 3, [1]),
])
def test_accumulate_pooling_split_columns(ctx, oracle, D, dtype, thr, k, counts):
    """Few label groups: the planner deals every group over several accumulator columns (plan B), the epilogue adds
    them up.  Same oracle, bit-exact, identical to the generic kernel."""
    rng = np.random.default_rng(D)
    case = synth.make_case(900 + D, counts, 180, D, rows_per_speaker=rng.choice([1, 2, 3], size=180), impostor_frac=0.2, neighbours=3)
    gpu = run_gpu(ctx, case, dtype, 0, thr, k, path=2, acc=2)
    path, nfb = ctx.last_path()
    assert path == 3, "expected the accumulate-pooling kernel"
    if thr > 0.3:
        assert nfb == 0
    assert_same(gpu, run_oracle(oracle, case, dtype, 0, thr, k), f"acc split D={D}")


@pytest.mark.parametrize("D,dtype,thr,k", [(192, 1, 0.354, 4), (64, 1, 0.354, 10), (256, 1, -1.0, 10), (320, 1, 0.354, 5),
                                           (512, 1, 0.2, 8), (192, 0, 0.354, 4), (100, 1, 0.354, 3)])
def test_accumulate_pooling_path(ctx, oracle, D, dtype, thr, k):
    """Mean pooling over >= 128 label groups runs the accumulate-pooling kernel (group-interleaved layout, pooled sum
    formed inside the MMA accumulation).  Same oracle, bit-exact; and identical to the generic tcgen05 kernel."""
    case = many_groups_case(500 + D, D, max_size=30 if D >= 320 else 60)
    gpu = run_gpu(ctx, case, dtype, 0, thr, k, path=2, acc=2)   # forced: this small ragged case pads beyond the auto limit
    path, nfb = ctx.last_path()
    assert path == 3, "expected the accumulate-pooling kernel"
    if thr > 0.3:
        assert nfb == 0
    ref = run_oracle(oracle, case, dtype, 0, thr, k)
    assert_same(gpu, ref, f"acc D={D}")
    gpu2 = run_gpu(ctx, case, dtype, 0, thr, k, path=2, acc=0)
    assert ctx.last_path()[0] == 2
    assert_same(gpu2, ref, f"generic D={D}")


@pytest.mark.parametrize("path,acc", [(1, 0), (2, 0), (2, 2)])
def test_host_pipeline_chunks(ctx, oracle, path, acc):
    """The host-buffer call cuts big batches at label boundaries and overlaps H2D with scoring; force many small
    chunks (1 MB) including empty labels at chunk edges and check the stitched result."""
    rng = np.random.default_rng(8)
    counts = [700, 0, 0, 900, 1500, 3, 0, 2200, 64, 1, 0]
    case = synth.make_case(88, counts, 300, 192, rows_per_speaker=rng.choice([1, 2], size=300), impostor_frac=0.2)
    ctx.set_option("chunk_mb", 1)
    try:
        gpu = run_gpu(ctx, case, 1, 0, 0.354, 6, path=path, acc=acc)
    finally:
        ctx.set_option("chunk_mb", 128)
    assert_same(gpu, run_oracle(oracle, case, 1, 0, 0.354, 6), "pipelined host path")


def test_zero_rows_and_scale_invariance(ctx, oracle):
    case = synth.config2(total=400, P=60)
    case.seg[5] = 0.0
    case.bank[7] = 0.0
    gpu = run_gpu(ctx, case, 0, 0, 0.354, 5, path=1)
    assert_same(gpu, run_oracle(oracle, case, 0, 0, 0.354, 5), "zero rows")


def test_assign_matches_oracle_and_python(ctx, oracle):
    from oracle import matching_np as mnp
    case = synth.config2(total=1000, P=200)
    # several rows per trust level, low threshold so that lists are long
    ctx.set_option("path", 1)
    ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=0)
    rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, threshold=0.05, k=8)
    for thr, min_trust in [(0.3, "low"), (0.1, "medium"), (0.0, "high"), (0.3, "none")]:
        ctx.assign(thr, min_trust)
        out = ctx.fetch(with_assign=True)
        code = _native.TRUST_CODES.get(min_trust, 99)
        a_idx, a_score, a_conf, c_idx, c_score = oracle.assign(out["row"], out["score"], out["trust"], out["count"], thr, code)
        assert np.array_equal(out["assign_idx"], a_idx)
        assert np.array_equal(out["assign_score"].view(np.uint64), a_score.view(np.uint64))
        assert np.array_equal(out["assign_conf"], a_conf)
        assert np.array_equal(out["cand_idx"], c_idx)
        assert np.array_equal(out["cand_score"].view(np.uint64), c_score.view(np.uint64))
        # and against the Python restatement of combine_signals fed the same per-label signal lists
        for g in range(case.G):
            sigs = []
            for i in range(out["count"][g]):
                trust = _native.TRUST_NAMES[out["trust"][g, i]]
                if not mnp.passes_min_trust(trust, min_trust):
                    continue
                sigs.append(mnp.Signal("embedding_match", str(int(out["row"][g, i])), float(out["score"][g, i]),
                                       {"trust_level": trust}))
            ref = mnp.combine_signals(f"S{g}", sigs, threshold=thr)
            got_id = None if out["assign_idx"][g] < 0 else str(int(out["row"][g, out["assign_idx"][g]]))
            assert got_id == ref["speaker_id"]
            assert out["assign_score"][g] == ref["score"]
            assert _native.CONF_NAMES[out["assign_conf"][g]] == ref["confidence"]
            got_c = [(str(int(out["row"][g, j])), out["cand_score"][g, n]) for n, j in enumerate(out["cand_idx"][g]) if j >= 0]
            assert got_c == [(c["speaker_id"], c["score"]) for c in ref["candidates"]]


@pytest.mark.parametrize("dtype,pool,path", [(1, 0, 1), (1, 1, 1), (0, 0, 1), (1, 0, 2), (1, 1, 2)])
def test_affinity_pooled(ctx, oracle, dtype, pool, path):
    case = synth.config5(N=1500, L=7, D=256)
    ctx.set_option("path", path)
    ctx.set_option("acc", 1)
    nl, ll = ctx.affinity_pooled(case.seg, case.seg_label, case.G, dtype=dtype, pool=pool)
    ref = oracle.affinity(case.seg, case.goff, mode=dtype, pool=pool)
    if path == 2:
        assert ctx.last_path()[0] == (3 if pool == 0 else 2), "mean pooling runs the accumulate-pooling kernel"
    if path == 1:
        assert np.array_equal(nl.view(np.uint32), ref.view(np.uint32))
    else:
        # tcgen05: same bf16 operands, exact products, fp32 accumulation in a different order
        np.testing.assert_allclose(nl, ref, rtol=0, atol=2e-5)
    q = np.rint(nl.astype(np.float64) * 2.0 ** 30).astype(np.int64)
    cnt = np.diff(case.goff)
    ll_ref = np.stack([q[case.goff[a]:case.goff[a + 1]].sum(axis=0) / (cnt[a] * 2.0 ** 30) for a in range(case.G)])
    assert np.array_equal(ll, ll_ref.astype(np.float32))
    # every segment is closest to its own label on planted data
    assert (nl.argmax(axis=1) == case.seg_label).mean() > 0.99


def test_affinity_pooled_skewed_labels_and_empty_label(ctx, oracle):
    """config-5 shape with Zipf-skewed labels and a label without segments (its column of the output is 0)."""
    counts = [1400, 0, 700, 350, 12, 1, 233, 90]
    case = synth.make_case(55, counts, 8, 256, impostor_frac=0.0)
    ctx.set_option("path", 2)
    ctx.set_option("acc", 1)
    nl, ll = ctx.affinity_pooled(case.seg, case.seg_label, case.G, dtype=1, pool=0)
    assert ctx.last_path()[0] == 3
    ref = oracle.affinity(case.seg, case.goff, mode=1, pool=0)
    np.testing.assert_allclose(nl, ref, rtol=0, atol=2e-5)
    assert (nl[:, 1] == 0).all()
    ctx.set_option("acc", 0)
    try:
        nl2, _ = ctx.affinity_pooled(case.seg, case.seg_label, case.G, dtype=1, pool=0)
        assert ctx.last_path()[0] == 2
    finally:
        ctx.set_option("acc", 1)
    np.testing.assert_allclose(nl2, ref, rtol=0, atol=2e-5)


def test_errors(ctx):
    case = synth.config1()
    ctx.set_option("path", 0)
    ctx.bank_load(case.bank, case.row_speaker, case.row_trust)
    with pytest.raises(_native.NativeError, match="non-decreasing"):
        ctx.identify(case.seg, case.seg_label[::-1].copy(), case.G)
    with pytest.raises(_native.NativeError, match="out of range"):
        ctx.identify(case.seg, case.seg_label + 5, case.G)
    with pytest.raises(_native.NativeError, match="k must be"):
        ctx.identify(case.seg, case.seg_label, case.G, k=33)
    with pytest.raises(_native.NativeError, match="contiguous"):
        ctx.bank_load(case.bank, np.array([0, 1, 0], dtype=np.int32))
    # empty input: no segments at all -> no matches
    ctx.bank_load(case.bank, case.row_speaker, case.row_trust)
    rows, scores, counts = ctx.identify(np.zeros((0, 192), np.float32), np.zeros(0, np.int32), 2)
    assert (counts == 0).all() and (rows == -1).all()
