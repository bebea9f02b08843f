"""fp16 segment embeddings (the compact sidecar form) through sdk_identify_f16 / sdk_identify_f16_dev: the result must
equal the oracle's on the SAME values widened to fp32 -- widening is exact, so everything downstream (canonical
normalise, scores, ids) is bit-identical.  Every K1 variant is covered: the vector kernel (64-bit loads), the scalar
kernel (D % 4 != 0 or a 2-byte-aligned pointer), the accumulate-pooling scatter form, and the chunked host pipeline."""
import numpy as np
import pytest

from speaker_diarization_toolkit_b200 import _native, synth
from test_gpu_parity import assert_same, ragged_case

pytestmark = pytest.mark.gpu


def _f16_case(case):
    h = case.seg.astype(np.float16)
    assert np.isfinite(h).all()
    return h, h.astype(np.float32)


@pytest.mark.parametrize("D,dtype,path,acc,pool", [(192, 1, 2, 2, 0), (192, 1, 2, 0, 0), (256, 0, 1, 0, 1), (100, 1, 2, 0, 0),
                                                   (512, 1, 2, 0, 1), (30, 0, 1, 0, 0)])
def test_f16_host_identify_matches_oracle_on_widened_input(ctx, oracle, D, dtype, path, acc, pool):
    case = ragged_case(3000 + D, D, P_speakers=200)
    h, wide = _f16_case(case)
    for k, v in (("path", path), ("acc", acc), ("gemv", 1), ("cand", 16), ("eps", -1.0)):
        ctx.set_option(k, v)
    ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=dtype)
    gpu = ctx.identify(h, case.seg_label, case.G, pool=pool, threshold=0.354, k=8)
    if acc == 2:
        assert ctx.last_path()[0] == 3
    ref = oracle.identify(wide, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=dtype, pool=pool, threshold=0.354, k=8)
    assert_same(gpu, ref, f"f16 D={D}")
    # and it is NOT simply the fp32 path's answer on the unrounded values unless the rounding happened to be harmless:
    # the fp32 entry point on the widened matrix gives the same bits as the fp16 entry point
    gpu32 = ctx.identify(wide, case.seg_label, case.G, pool=pool, threshold=0.354, k=8)
    assert_same(gpu32, ref, f"f32-on-widened D={D}")


def test_f16_chunked_pipeline(ctx, oracle):
    """the host pipeline cuts chunks in BYTES of the stored type: small chunk_mb forces several chunks of fp16 rows"""
    rng = np.random.default_rng(8)
    case = synth.make_case(3100, synth.zipf_counts(rng, 9000, 40), 300, 192, rows_per_speaker=rng.choice([1, 2], size=300), impostor_frac=0.2)
    h, wide = _f16_case(case)
    for k, v in (("path", 0), ("acc", 1), ("gemv", 1), ("cand", 16), ("eps", -1.0), ("chunk_mb", 1)):
        ctx.set_option(k, v)
    try:
        ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=1)
        gpu = ctx.identify(h, case.seg_label, case.G, pool=0, threshold=0.354, k=5)
    finally:
        ctx.set_option("chunk_mb", 128)
    ref = oracle.identify(wide, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=1, pool=0, threshold=0.354, k=5)
    assert_same(gpu, ref, "f16 chunked")


@pytest.mark.parametrize("offset_elems", [0, 1, 4])
def test_f16_device_pointers_and_alignment(ctx, oracle, offset_elems):
    """_dev entry point with a tensor view at an element offset: 2-byte aligned pointers take the scalar kernel, 8-byte
    aligned ones the vector / scatter kernels; same bits either way."""
    torch = pytest.importorskip("torch")
    case = ragged_case(3200, 192, P_speakers=150)
    h, wide = _f16_case(case)
    N = h.shape[0]
    flat = torch.zeros(N * 192 + 16, dtype=torch.float16, device="cuda")
    view = flat[offset_elems:offset_elems + N * 192]
    view.copy_(torch.from_numpy(h.reshape(-1)))
    lab = torch.from_numpy(case.seg_label.astype(np.int32)).cuda()
    for k, v in (("path", 2), ("acc", 1), ("gemv", 1), ("cand", 16), ("eps", -1.0)):
        ctx.set_option(k, v)
    ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=1)
    torch.cuda.synchronize()
    ctx.identify_dev(view.data_ptr(), lab.data_ptr(), N, case.G, 0, 0.354, 6, f16=True)
    out = ctx.fetch()
    ref = oracle.identify(wide, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=1, pool=0, threshold=0.354, k=6)
    assert_same((out["row"], out["score"], out["count"]), ref, f"f16 dev offset {offset_elems}")
