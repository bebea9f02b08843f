"""GPU tests of the device-pointer entry points (`sdk_identify_dev`, `sdk_affinity_pooled_dev`, `sdk_bank_load_dev`) and of
the list merge K4 (`sdk_merge_topk`, the kernel that follows the all-gather) on ONE GPU.  torch only supplies device memory."""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native, sharding, synth

pytestmark = pytest.mark.gpu


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _identify_dev(torch, ctx, case, dtype, pool, thr, k, offset_elems):
    """Uploads the case; the segment matrix starts `offset_elems` floats into its allocation (a tensor view)."""
    bank, spk, tr = _dev(torch, case.bank), _dev(torch, case.row_speaker), _dev(torch, case.row_trust)
    P, D = case.bank.shape
    ctx.bank_load_dev(bank.data_ptr(), spk.data_ptr(), tr.data_ptr(), P, D, dtype)
    N = case.seg.shape[0]
    buf = torch.zeros(N * D + 64, dtype=torch.float32, device="cuda")
    seg = buf[offset_elems:offset_elems + N * D]
    seg.copy_(_dev(torch, case.seg).reshape(-1))
    assert seg.data_ptr() % 16 == (4 * offset_elems) % 16
    lab = _dev(torch, case.seg_label)
    torch.cuda.synchronize()
    ctx.identify_dev(seg.data_ptr(), lab.data_ptr(), N, case.G, pool, thr, k)
    out = ctx.fetch()
    return out["row"], out["score"], out["count"]


@pytest.mark.parametrize("offset", [0, 1, 2, 3])
@pytest.mark.parametrize("acc", [1, 2])
def test_identify_dev_any_float_alignment(ctx, oracle, offset, acc):
    """a legal 4-byte aligned float* (a view at an element offset) must not reach the 128-bit load paths.
    Accumulate-pooling shape (mean, D = 192, many labels): aligned pointers take path 3, the others the generic kernel."""
    torch = pytest.importorskip("torch")
    case = synth.config3(recordings=2, seg_per_rec=700, P=900, D=192)
    ctx.set_option("path", 2)
    ctx.set_option("acc", acc)
    ctx.set_option("gemv", 1)
    ctx.set_option("cand", 16)
    ctx.set_option("eps", -1.0)
    got = _identify_dev(torch, ctx, case, 1, 0, 0.354, 4, offset)
    ref = oracle.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=1, pool=0, threshold=0.354, k=4)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2])
    assert np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32))
    assert ctx.last_path()[0] == (3 if offset == 0 else 2)
    ctx.set_option("path", 0)
    ctx.set_option("acc", 1)


@pytest.mark.parametrize("offset", [0, 1])
def test_affinity_dev_any_float_alignment(ctx, oracle, offset):
    torch = pytest.importorskip("torch")
    case = synth.config5(N=3000, L=16, D=256)
    N, D = case.seg.shape
    buf = torch.zeros(N * D + 64, dtype=torch.float32, device="cuda")
    seg = buf[offset:offset + N * D]
    seg.copy_(_dev(torch, case.seg).reshape(-1))
    lab = _dev(torch, case.seg_label)
    nl = torch.empty((N, case.G), dtype=torch.float32, device="cuda")
    ctx.set_option("path", 2)
    ctx.set_option("acc", 1)
    ctx.affinity_pooled_dev(seg.data_ptr(), lab.data_ptr(), N, D, case.G, 1, 0, nl.data_ptr(), None)
    ctx.sync()
    ref = oracle.affinity(case.seg, case.goff, mode=1, pool=0)
    np.testing.assert_allclose(nl.cpu().numpy(), ref, rtol=0, atol=2e-3)          # stage-A arithmetic (bf16 operands, fp32 accumulate)
    ctx.set_option("path", 0)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("dtype,pool,thr,k", [(1, 0, -1.0, 10), (0, 1, 0.3, 5)])
def test_merge_topk_of_shard_lists_equals_whole_bank(ctx, oracle, world, dtype, pool, thr, k):
    """K4 on one GPU: the bank is cut on speaker boundaries into `world` shards, every shard is identified on its own
    (global row offsets), and sdk_merge_topk of the per-shard lists must equal the oracle on the whole bank -- rows,
    order, scores bit for bit, trust codes, and the assignment computed from the merged lists."""
    rng = np.random.default_rng(world)
    rps = rng.choice([1, 2, 3], size=300)
    case = synth.make_case(77 + world, synth.zipf_counts(rng, 500, 6), 300, 128, rows_per_speaker=rps, neighbours=5, impostor_frac=0.1)
    shards = sharding.shard_bank_rows(case.row_speaker, world)
    ctx.set_option("path", 0)
    lists = []
    for p0, p1 in shards:
        assert p1 > p0
        ctx.bank_load(case.bank[p0:p1], case.row_speaker[p0:p1], case.row_trust[p0:p1], dtype=dtype, global_row_offset=p0)
        ctx.identify(case.seg, case.seg_label, case.G, pool=pool, threshold=thr, k=k)
        lists.append(ctx.fetch())
    ctx.merge_topk(np.stack([o["row"] for o in lists]), np.stack([o["score"] for o in lists]),
                   np.stack([o["count"] for o in lists]), np.stack([o["trust"] for o in lists]))
    ctx.assign(0.2, "low")
    out = ctx.fetch(with_assign=True)
    ref = oracle.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=dtype, pool=pool, threshold=thr, k=k)
    assert np.array_equal(out["row"], ref[0]) and np.array_equal(out["count"], ref[2])
    assert np.array_equal(out["score"].view(np.uint32), ref[1].view(np.uint32))
    tr_ref = np.where(ref[0] >= 0, case.row_trust[np.clip(ref[0], 0, None)], 4).astype(np.uint8)
    assert np.array_equal(out["trust"], tr_ref)
    a = oracle.assign(ref[0], ref[1], tr_ref, ref[2], 0.2, 2)
    assert np.array_equal(out["assign_idx"], a[0]) and np.array_equal(out["assign_score"], a[1])
    # host mirror of the kernel agrees as well
    hr, hs, hc, _ = sharding.merge_topk_lists(np.stack([o["row"] for o in lists]), np.stack([o["score"] for o in lists]),
                                              np.stack([o["count"] for o in lists]), k)
    assert np.array_equal(hr, out["row"]) and np.array_equal(hc, out["count"])


def test_merge_topk_ties_and_padding(ctx):
    """Equal scores across shards order by global row; short and empty lists are padded with -1."""
    rows = np.full((3, 2, 4), -1, np.int64)
    scores = np.zeros((3, 2, 4), np.float32)
    counts = np.zeros((3, 2), np.int32)
    rows[0, 0, :2], scores[0, 0, :2], counts[0, 0] = [40, 7], [0.5, 0.25], 2
    rows[1, 0, :3], scores[1, 0, :3], counts[1, 0] = [12, 90, 91], [0.5, 0.5, 0.1], 3
    rows[2, 0, :1], scores[2, 0, :1], counts[2, 0] = [3], [0.25], 1
    ctx.merge_topk(rows, scores, counts)
    out = ctx.fetch()
    assert out["row"][0].tolist() == [12, 40, 90, 3] and out["count"].tolist() == [4, 0]
    assert out["score"][0].tolist() == [0.5, 0.5, 0.5, 0.25]
    assert out["row"][1].tolist() == [-1, -1, -1, -1]
    with pytest.raises(_native.NativeError):
        ctx.merge_topk(rows, scores, np.full((3, 2), 9, np.int32))


def test_empty_shard_is_rejected_without_a_world(ctx):
    with pytest.raises(_native.NativeError):
        ctx.bank_load(np.zeros((0, 64), np.float32), np.zeros(0, np.int32))
