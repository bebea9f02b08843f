"""Phase trace of the small-query kernels (k_gemv8 / k_gemv8_tail) on config 4-i.

  python tools/gv_trace.py --build     here: compiles gemv.cu with -DSDK_GV_TRACE and links libsdk_b200_trace.so
  python tools/gv_trace.py [--lib X]   on the GPU box: runs 4-i through that library and prints the phase table

The trace build is a diagnostic; the shipped libsdk_b200.so holds none of it."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
PKG = ROOT / "speaker_diarization_toolkit_b200"
TRACE_SO = PKG / "libsdk_b200_trace.so"


def _arg(name, default=None, many=False):
    vals = [sys.argv[i + 1] for i, a in enumerate(sys.argv[:-1]) if a == name]
    return vals if many else (vals[-1] if vals else default)


def build():
    """--build [--define NAME[=V]]... [--out libsdk_b200_traceX.so]"""
    from speaker_diarization_toolkit_b200 import build as b
    b.build()
    out = PKG / _arg("--out", TRACE_SO.name)
    obj = PKG / "build" / (out.stem + "_gemv.o")
    defs = ["-D" + d for d in _arg("--define", many=True)]
    subprocess.run([b.nvcc(), *b.NVCC_FLAGS, "-DSDK_GV_TRACE", *defs, "-c", str(b.CSRC / "gemv.cu"), "-o", str(obj)], check=True)
    objs = [str(PKG / "build" / (s + ".o")) for s in b.SOURCES if s != "gemv.cu"] + [str(obj)]
    subprocess.run([b.nvcc(), "-shared", "-o", str(out), *objs, "-cudart", "shared", "-ldl"], check=True)
    print(out)


def main():
    os.environ["SDK_B200_LIB"] = str(PKG / _arg("--lib", TRACE_SO.name))
    print("library", os.environ["SDK_B200_LIB"])
    import ctypes as C
    import numpy as np
    import torch
    from speaker_diarization_toolkit_b200 import _native
    lib = _native.load()
    P, D, N = 125000, 512, 8
    g = torch.Generator(device="cuda").manual_seed(404)
    bank = torch.randn(P, D, device="cuda", generator=g)
    spk = torch.arange(P, device="cuda", dtype=torch.int32)
    seg = torch.randn(N, D, device="cuda", generator=g)
    lab = torch.arange(N, device="cuda", dtype=torch.int32)
    ctx = _native.Context(0)
    ctx.bank_load_dev(bank.data_ptr(), spk.data_ptr(), None, P, D, _native.DTYPE_BF16)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = np.zeros((2, 160, 24), dtype=np.uint64)
    rows = []
    for it in range(8):
        flush.fill_(it)
        flush.sum()
        torch.cuda.synchronize()
        ctx.identify_dev(seg.data_ptr(), lab.data_ptr(), N, N, 0, -1.0, 10)
        torch.cuda.synchronize()
        assert ctx.last_path()[0] == 4, ctx.last_path()
        rc = lib.sdk_debug_gv_trace(out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        if it >= 3:
            rows.append(out.copy())
    names0 = {1: "start", 2: "queries normalised", 3: "B fragments", 4: "pass0 mma (warp 0)", 5: "pass0 mma (all)", 6: "pass0 select (warp 0)",
              7: "pass0 select (all)", 8: "pass1 mma (warp 0)", 9: "pass1 mma (all)", 10: "pass1 select (warp 0)", 11: "pass1 select (all)", 16: "published"}
    names1 = {2: "after wait", 3: "lists in smem", 4: "T0", 5: "compaction", 6: "rank", 7: "bound", 8: "staged chunk 0", 9: "stage B", 10: "select"}
    for kern, names, ncta in ((0, names0, 148), (1, names1, 8)):
        print("kernel", "k_gemv8" if kern == 0 else "k_gemv8_tail")
        t = np.stack([r[kern, :ncta] for r in rows]).astype(np.int64)          # [it, cta, slot]
        gt0, gt1 = (0, 17) if kern == 0 else (1, 11)
        c0, c1 = (1, 16) if kern == 0 else (2, 10)
        span_ns = (t[:, :, gt1].max(1) - t[:, :, gt0].min(1))
        print("  first start -> last end (globaltimer): %s ns" % span_ns.tolist())
        if kern == 1:
            print("  launch -> wait released (globaltimer): %s ns" % (t[:, :, 1] - t[:, :, 0]).mean(1).round().tolist())
        ghz = ((t[:, :, c1] - t[:, :, c0]) / np.maximum(1, (t[:, :, gt1] - t[:, :, gt0 if kern == 0 else 1]))).mean()
        print("  SM clock ~ %.2f GHz" % ghz)
        start_skew = (t[:, :, gt0] - t[:, :, gt0].min(1, keepdims=True))
        print("  CTA start skew ns: mean %.0f max %.0f" % (start_skew.mean(), start_skew.max()))
        prev = c0
        for slot in sorted(names):
            if slot == c0:
                continue
            d = (t[:, :, slot] - t[:, :, prev]) / ghz / 1000.0                 # us
            cum = (t[:, :, slot] - t[:, :, c0]) / ghz / 1000.0
            print("  %-24s +%6.2f us (max CTA %6.2f)   cumulative %6.2f (max %6.2f)" % (names[slot], d.mean(), d.max(1).mean(), cum.mean(), cum.max(1).mean()))
            prev = slot
        if kern == 0:
            us = lambda a, b: ((t[:, :, a] - t[:, :, b]) / ghz / 1000.0).mean()
            print("  prep detail: query loads arrived +%.2f us after start, norm math +%.2f us" % (us(18, 1), us(19, 18)))
            print("  last selection (warp 0): keys +%.2f us after the sync, T0 +%.2f, compaction +%.2f, rank/scatter +%.2f; survivors m = %.1f (max %d)"
                  % (us(20, 9), us(21, 20), us(22, 21), us(10, 22), t[:, :, 12].mean(), t[:, :, 12].max()))
    ctx.close()


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        main()
