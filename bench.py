#!/usr/bin/env python3
"""bench.py -- scored segment x profile pairs / second on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg3|cfg1|cfg2|cfg4|cfg4i|cfg4ii|cfg5] [--pool mean|max]

Default workload = BASELINE.json configs[2] ("cfg3": batch of 10k recordings, ~20M segments, vs a 10k-profile
bank, 192-d, bf16 operands) -- the configuration the metric "pairs/sec at 1/2/4/8 B200" is quoted on; it fits one
GPU.  For N > 1 the recordings are split over the ranks (independent units, NO data-path collective), total
work fixed -> "scaling": "strong".  A step = one pass of the hot path (normalise -> tcgen05 pooled GEMM ->
canonical re-score -> top-k -> assignment) over the rank's batch.
  value       inputs resident in HBM (device pointers through the C-ABI), CUDA-event timed on the library stream
  e2e         same metric through the host-buffer C-ABI call: pinned host -> device copies and result read-back
              inside the timed region
  roofline    dominant kernel (tcgen05 pooled GEMM) against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the NumPy/OpenBLAS port of the path (oracle/matching_np.py) on a bounded sample, all host cores
`--impl reference` times that CPU port only (rank 0), same metric/config.
Synthetic data (the reference's embedding extractors are network APIs): planted speakers, sigma 0.35, 10 % impostors.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path


def _host_threads_for_the_cpu_arm():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference arm (rank 0 only, the other ranks exit)
    must time the CPU path with ALL host threads, so the BLAS thread count is set before NumPy is imported."""
    if "--impl" in sys.argv and "reference" in sys.argv[sys.argv.index("--impl") + 1:sys.argv.index("--impl") + 2]:
        n = str(os.cpu_count() or 1)
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
            os.environ[v] = n


_host_threads_for_the_cpu_arm()

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "scored segment x profile pairs per second"
UNIT = "pairs/s"

WORKLOADS = {
    "cfg1": dict(R=1, seg=40, labels=2, P=3, D=192, dtype=0, k=3, thr=0.354, seed=101,
                 desc="configs[0]: speaker-assign CLI, 2 labels / 40 segments vs 3 enrolled profiles, 192-d fp32 (process-spawn bound)"),
    # name: recordings, segments/recording (Poisson mean), labels/recording, bank rows, D, dtype, k, threshold
    "cfg3": dict(R=10000, seg=2000, labels=8, P=10000, D=192, dtype=1, k=4, thr=0.354, seed=303,
                 desc="configs[2]: 10k recordings (~20M segments) vs 10k-profile bank, 192-d, bf16 operands, DP over recordings"),
    "cfg2": dict(R=1, seg=2000, labels=8, P=500, D=256, dtype=0, k=10, thr=0.354, seed=202,
                 desc="configs[1]: 1-hour meeting, 8 labels x 2k segments vs 500-profile bank, 256-d fp32 (latency-bound)"),
    "cfg5": dict(R=1, seg=50000, labels=16, P=50000, D=256, dtype=1, k=0, thr=0.0, seed=505,
                 desc="configs[4]: 50k x 50k segment-segment cosine affinity pooled per label (16 labels), 256-d bf16, single GPU"),
    "cfg4": dict(R=64, seg=2000, labels=8, P=125000, D=512, dtype=1, k=10, thr=-1.0, seed=404,
                 desc="configs[3]: 125k bank rows PER GPU (1M at 8 GPUs), 512-d bf16, top-10 per label, NCCL all-gather merge"),
    # the other two query shapes SURVEY 8d names for the sharded bank: (i) 8 pooled label centroids (bank stream,
    # HBM-bound), (ii) one meeting's 2 000 raw segments
    "cfg4i": dict(R=1, seg=8, labels=8, P=125000, D=512, dtype=1, k=10, thr=-1.0, seed=404,
                  desc="configs[3] variant (i): 8 label centroids vs 125k bank rows PER GPU, 512-d bf16, top-10 (HBM-bound bank stream)"),
    "cfg4ii": dict(R=1, seg=2000, labels=8, P=125000, D=512, dtype=1, k=10, thr=-1.0, seed=404,
                   desc="configs[3] variant (ii): one meeting (8 labels, ~2k segments) vs 125k bank rows PER GPU, 512-d bf16, top-10"),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(workload, kernel_label, same_launch_shape):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this very workload
    (tools/ncu_summary.py --traffic-key); None when the launch being timed is not the one that was captured."""
    tfile = ROOT / "profiles" / "roofline_traffic.json"
    if not same_launch_shape or not tfile.exists():
        return None
    ent = json.loads(tfile.read_text()).get(workload)
    if not ent or ent.get("kernel", "?") not in kernel_label:
        return None
    return ent["dram_bytes_per_launch"]


# ---- synthetic batch (group counts on the host, embeddings generated on the device) -------------------
def group_counts(cfg, scale):
    rng = np.random.default_rng(cfg["seed"])
    R = max(1, int(round(cfg["R"] * scale)))
    L = cfg["labels"]
    w = 1.0 / np.arange(1, L + 1)
    counts = np.empty((R, L), dtype=np.int64)
    tot = np.maximum(L, rng.poisson(cfg["seg"], size=R))
    for r in range(R):
        ww = rng.permutation(w)
        c = np.maximum(1, np.floor(ww / ww.sum() * tot[r]).astype(np.int64))
        c[np.argmax(c)] += tot[r] - c.sum()
        counts[r] = c
    truth = rng.integers(0, cfg["P"], size=(R, L))
    truth[rng.random((R, L)) < 0.1] = -1
    return counts, truth


def make_bank(torch, cfg, dev, n_rows, row0=0):
    g = torch.Generator(device=dev)
    g.manual_seed(cfg["seed"] * 7919 + 1)
    D = cfg["D"]
    # centroids are generated for the whole bank id space so every rank agrees on them
    cent = torch.randn((cfg["P_total"], D), generator=g, device=dev, dtype=torch.float32)
    cent = cent / cent.norm(dim=1, keepdim=True)
    rows = cent[row0:row0 + n_rows] + 0.35 / math.sqrt(D) * torch.randn((n_rows, D), generator=g, device=dev)
    rows = rows / rows.norm(dim=1, keepdim=True) * (0.5 + 19.5 * torch.rand((n_rows, 1), generator=g, device=dev))
    return cent, rows.contiguous()


def make_segments(torch, cfg, dev, cent, counts, truth, seed):
    """counts/truth: [R_local, L].  Returns (seg [N,D] fp32 device, label [N] int32 device, N, G)."""
    D = cfg["D"]
    flat_c = counts.reshape(-1)
    G = flat_c.shape[0]
    N = int(flat_c.sum())
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lab = torch.repeat_interleave(torch.arange(G, device=dev, dtype=torch.int32), torch.as_tensor(flat_c, device=dev))
    t = torch.as_tensor(truth.reshape(-1), device=dev)
    imp = torch.randn((G, D), generator=g, device=dev)
    imp = imp / imp.norm(dim=1, keepdim=True)
    gcent = torch.where((t >= 0)[:, None], cent[t.clamp(min=0)], imp)           # [G, D]
    seg = torch.empty((N, D), device=dev, dtype=torch.float32)
    step = 1 << 21
    for a in range(0, N, step):
        b = min(N, a + step)
        x = gcent[lab[a:b].long()] + 0.35 / math.sqrt(D) * torch.randn((b - a, D), generator=g, device=dev)
        x = x / x.norm(dim=1, keepdim=True) * (0.5 + 19.5 * torch.rand((b - a, 1), generator=g, device=dev))
        seg[a:b] = x
    return seg, lab, N, G


# ---- NUMA placement of the pinned host buffers ---------------------------------------------------------------------
_ALL_CPUS = None


def _node_cpus(node):
    cpus = set()
    for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
        a, _, b = part.partition("-")
        if a:
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(torch, local_rank):
    """Pinned host memory is placed by first touch: bind this process to the CPUs of the NUMA node its GPU hangs off
    before allocating, so that with 8 ranks the H2D streams read from both sockets' memory instead of one.  The node comes
    from sysfs; where sysfs does not say (-1: virtualised PCI topology) and the host has several nodes, every node is tried
    with a 64 MB pinned buffer and the one with the fastest host->device copy wins.  Returns (node, how) or (None, why)."""
    global _ALL_CPUS
    try:
        if _ALL_CPUS is None:
            _ALL_CPUS = os.sched_getaffinity(0)
        nodes = sorted(int(p.name[4:]) for p in Path("/sys/devices/system/node").glob("node[0-9]*"))
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = -1
        try:
            node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        except Exception:
            node = -1
        how = "sysfs"
        if node < 0:
            if len(nodes) < 2:
                return None, "single NUMA node"
            dev = torch.device("cuda", local_rank)
            best, rates = None, {}
            d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
            for nd in nodes:
                cpus = _node_cpus(nd) & _ALL_CPUS
                if not cpus:
                    continue
                os.sched_setaffinity(0, cpus)
                h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
                h.fill_(1)
                d.copy_(h, non_blocking=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4):
                    d.copy_(h, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                rates[nd] = 4 * h.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
                del h
                if best is None or rates[nd] > rates[best]:
                    best = nd
            os.sched_setaffinity(0, _ALL_CPUS)
            if best is None:
                return None, "no usable NUMA node"
            node, how = best, "h2d probe " + ", ".join(f"node{k}: {v:.1f} GB/s" for k, v in sorted(rates.items()))
        cpus = _node_cpus(node) & _ALL_CPUS
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node, how
    except Exception as exc:
        unbind_cpus()                                       # (a probe that failed half way must not leave the process bound)
        return None, f"unavailable ({type(exc).__name__})"


def unbind_cpus():
    if _ALL_CPUS:
        try:
            os.sched_setaffinity(0, _ALL_CPUS)
        except Exception:
            pass


# ---- clocks sampler ------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    """nvidia-smi sampler.  Started BEFORE the warm-up (the tool needs ~0.2 s to produce its first line), every sample is
    stamped on arrival, and only the samples that fall inside the timed region [mark_begin, stop] are reported; a region
    shorter than the sampling period falls back to the samples of the last half second under load."""

    def __init__(self, dev_index):
        self.samples, self.proc, self.dev = [], None, dev_index
        self.t_begin = None

    def mark_begin(self):
        self.t_begin = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_wait = time.time() + 2.0            # short workloads finish before the tool prints its first line: wait for it
            while not self.samples and time.time() < t_wait:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        t_end = time.time()
        time.sleep(0.08)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t_begin if self.t_begin is not None else 0.0
        inside = [ln for ts, ln in self.samples if t0 <= ts <= t_end + 0.03]
        if len(inside) < 2:
            inside = [ln for ts, ln in self.samples if t_end - 0.5 <= ts <= t_end + 0.03]
        for s in inside:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                mhz.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


# ---- CPU port (cpu_baseline and --impl reference) -------------------------------------------------
def blas_threads():
    """threads the BLAS behind NumPy will really use (threadpoolctl), else the environment's word for it"""
    try:
        from threadpoolctl import threadpool_info
        n = [int(i.get("num_threads", 0)) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return int(os.environ.get("OMP_NUM_THREADS") or os.cpu_count() or 1)


def cpu_port_rate(cfg, counts, truth, budget_s, rec_cap=None):
    """Times oracle/matching_np.identify (+ the combine_signals restatement) recording by recording on the host.
    Returns (pairs/s, recordings timed, seconds)."""
    from oracle import matching_np as mnp
    from speaker_diarization_toolkit_b200 import synth
    P, D, L = cfg["P_cpu"], cfg["D"], cfg["labels"]
    rng = np.random.default_rng(cfg["seed"] + 1)
    cent = rng.standard_normal((P, D)).astype(np.float32)
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    bank = cent + (0.35 / math.sqrt(D)) * rng.standard_normal((P, D)).astype(np.float32)
    row_speaker = np.arange(P, dtype=np.int32)
    trust_names = ["high", "medium", "low"]
    pairs, t_used, n_rec = 0, 0.0, 0
    R = counts.shape[0]
    while t_used < budget_s and n_rec < (rec_cap or R):
        r = n_rec % R
        c = counts[r]
        goff = np.r_[0, np.cumsum(c)]
        n = int(goff[-1])
        tr = np.repeat(np.where(truth[r] >= 0, truth[r] % P, 0), c)
        seg = cent[tr] + (0.35 / math.sqrt(D)) * rng.standard_normal((n, D)).astype(np.float32)
        t0 = time.perf_counter()
        rows, scores, cnt = mnp.identify(seg, goff, bank, row_speaker, mode=cfg["dtype"], pool=cfg.get("pool", 0), threshold=cfg["thr"], k=cfg["k"])
        for g in range(L):
            sigs = [mnp.Signal("embedding_match", str(int(rows[g, i])), float(scores[g, i]),
                               {"trust_level": trust_names[int(rows[g, i]) % 3]}) for i in range(cnt[g])]
            mnp.combine_signals(f"S{g}", sigs, threshold=0.3)
        t_used += time.perf_counter() - t0
        pairs += n * P
        n_rec += 1
    return pairs / max(t_used, 1e-9), n_rec, t_used


def run_reference(args, cfg, counts, truth):
    """`--impl reference`: the CPU port on the host cores, all of them; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:                                    # in case a BLAS was initialised with the launcher's OMP_NUM_THREADS=1 anyway
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    nthreads = blas_threads()
    per_step = max(1.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_port_rate(cfg, counts, truth, per_step)
    tot_pairs, tot_t, tot_rec = 0.0, 0.0, 0
    for _ in range(args.steps):
        rate, n_rec, t = cpu_port_rate(cfg, counts, truth, per_step)
        tot_pairs += rate * t
        tot_t += t
        tot_rec += n_rec
    value = tot_pairs / tot_t
    sample = f"{tot_rec / max(1, args.steps):.0f} recordings per step (~{per_step:.0f} s of NumPy/OpenBLAS on {nthreads} threads), bank {cfg['P_cpu']} rows"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "pool": "max" if cfg.get("pool") else "mean"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "host_cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---- config 5: pooled self-affinity (single GPU; the path does not need a collective) -------------------------
def run_cfg5(args, cfg):
    import torch
    from speaker_diarization_toolkit_b200 import _native
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    N, L, D = max(1024, int(cfg["seg"] * args.scale)), cfg["labels"], cfg["D"]
    rng = np.random.default_rng(cfg["seed"])
    w = rng.permutation(1.0 / np.arange(1, L + 1))
    cnt = np.maximum(1, np.floor(w / w.sum() * N).astype(np.int64))
    cnt[np.argmax(cnt)] += N - cnt.sum()
    g = torch.Generator(device=dev)
    g.manual_seed(cfg["seed"])
    cent = torch.randn((L, D), generator=g, device=dev)
    cent = cent / cent.norm(dim=1, keepdim=True)
    lab = torch.repeat_interleave(torch.arange(L, device=dev, dtype=torch.int32), torch.as_tensor(cnt, device=dev))
    seg = cent[lab.long()] + 0.35 / math.sqrt(D) * torch.randn((N, D), generator=g, device=dev)
    seg = (seg / seg.norm(dim=1, keepdim=True) * (0.5 + 19.5 * torch.rand((N, 1), generator=g, device=dev))).contiguous()
    out_nl = torch.empty((N, L), device=dev, dtype=torch.float32)
    out_ll = torch.empty((L, L), device=dev, dtype=torch.float32)
    ctx = _native.Context(0)
    ctx.set_option("profile", 1)
    ctx.set_option("cta_group", args.cta_group)
    ctx.set_option("acc", args.acc)

    def step():
        ctx.affinity_pooled_dev(seg.data_ptr(), lab.data_ptr(), N, D, L, 1, 0, out_nl.data_ptr(), out_ll.data_ptr())

    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)          # > L2: inputs (51 MB) would otherwise stay cached
    clocks = Clocks(0)
    clocks.start()
    for _ in range(args.warmup):
        step()
    ctx.sync()
    ctx.profile_reset()
    l0 = ctx.launch_count()
    clocks.mark_begin()
    tot = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.timer_start()
        step()
        tot += ctx.timer_stop()
    clk = clocks.stop()
    launches = ctx.launch_count() - l0
    pairs = float(N) * float(N)
    value = pairs * args.steps / (tot * 1e-3)
    pk, pk_src = peaks()
    gms, gl = ctx.profile_get("poolgemm")
    avg_ms = gms / max(1, gl)
    ach = 2.0 * pairs * D / (avg_ms * 1e-3) / 1e12
    kname = "k_poolacc (tcgen05, label columns split, mean pooling inside the MMA accumulation, dense output)" if ctx.last_path()[0] == 3 \
        else f"k_poolgemm (tcgen05 cta_group::{args.cta_group}, dense pooled output)"
    roof = {"kernel": kname, "bound": "tensor", "achieved": ach,
            "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
            "traffic": ncu_traffic("cfg5", kname, args.scale == 1.0),
            "peak_source": f"{pk_src} bf16 burst (kernel timed alone, ~1 ms)", "avg_launch_ms": avg_ms}
    # e2e through the host-pointer C-ABI call: inputs in PINNED host memory (the contract's rule), H2D + D2H inside
    h_seg = torch.empty((N, D), dtype=torch.float32, pin_memory=True)
    h_lab = torch.empty((N,), dtype=torch.int32, pin_memory=True)
    h_seg.copy_(seg)
    h_lab.copy_(lab)
    torch.cuda.synchronize()
    hs, hl = h_seg.numpy(), h_lab.numpy()
    ctx.affinity_pooled(hs, hl, L, dtype=1, pool=0)
    t0 = time.perf_counter()
    e_steps = max(1, min(3, args.steps))
    for _ in range(e_steps):
        nl, ll = ctx.affinity_pooled(hs, hl, L, dtype=1, pool=0)
    te = time.perf_counter() - t0
    e2e = {"value": pairs * e_steps / te, "unit": UNIT, "h2d_bytes_per_step": int(hs.nbytes + hl.nbytes),
           "d2h_bytes_per_step": int(nl.nbytes + ll.nbytes), "steps": e_steps}
    cpu = None
    if not args.no_cpu:
        from oracle import matching_np as mnp
        n_c = min(N, 8192)
        xs = mnp.bf16_round(mnp.l2_normalize(hs))
        t0 = time.perf_counter()
        S = xs[:n_c] @ xs.T
        starts = np.r_[0, np.cumsum(cnt)[:-1]]
        pooled = np.add.reduceat(S, starts, axis=1) / cnt[None, :].astype(np.float32)
        tc = time.perf_counter() - t0
        cpu = {"value": float(n_c) * N / tc, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {n_c} rows of the affinity ({tc:.1f} s, NumPy/OpenBLAS), extrapolates linearly"}
        del S, pooled
    print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": tot / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": cfg["desc"], "segments": N, "labels": L, "dim": D, "pool": "mean",
                                 "l2": "256 MB flush buffer written between timed iterations", "outputs": "[N,L] and [L,L] fp32"},
                      "clocks": clk, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e}))
    ctx.close()
    return 0


# ---- config 1: the product-level entry, through the real CLIs (process spawn is the reference's true cost here) ----
def run_cfg1(args):
    """BASELINE.json configs[0]: `speaker-assign assign AUDIO -t TRANSCRIPT --use-embeddings` on one synthetic
    Speechmatics-format transcript (2 labels, 40 segments) vs 3 enrolled profiles, 192-d ("cfg 1 timed through
    the actual CLI ... because the reference's real cost there is process spawn", speaker-assign:283-294).
      value  the in-process hot path on device-resident inputs (sdk_identify_dev + sdk_assign, CUDA events)
      e2e    wall time of the whole `bin/speaker-assign` process (interpreter start, CUDA context, store + sidecar reads,
             the backend plugin, H2D / D2H, JSON out), one process per step as the reference CLI does it
    The reference's own CLI chain cannot run on this box (/root/reference does not travel), and the chain spawns one
    `speaker_detection identify` process PER LABEL (speaker-assign:545-563) where this build makes one in-process call."""
    import tempfile
    import torch
    from oracle import canonical, matching_np as mnp
    from speaker_diarization_toolkit_b200 import _native, synth
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    case = synth.config1()
    N, P, D, L, k = case.seg.shape[0], case.bank.shape[0], case.seg.shape[1], case.G, 3
    pairs = float(N) * P
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = _native.Context(0)
    ctx.set_option("profile", 1)
    ctx.bank_load(case.bank, case.row_speaker, case.row_trust, dtype=_native.DTYPE_F32)
    seg = torch.from_numpy(case.seg).to(dev)
    lab = torch.from_numpy(case.seg_label.astype(np.int32)).to(dev)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    clocks = Clocks(0)
    clocks.start()
    for _ in range(max(3, args.warmup)):
        ctx.identify_dev(seg.data_ptr(), lab.data_ptr(), N, L, 0, 0.354, k)
        ctx.assign(0.3, "low")
    ctx.sync()
    ctx.profile_reset()
    l0 = ctx.launch_count()
    clocks.mark_begin()
    steps = max(args.steps, 20)
    ms = 0.0
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.timer_start()
        ctx.identify_dev(seg.data_ptr(), lab.data_ptr(), N, L, 0, 0.354, k)
        ctx.assign(0.3, "low")
        ms += ctx.timer_stop()
    clk = clocks.stop()
    launches = ctx.launch_count() - l0
    out = ctx.fetch(with_assign=True)
    ref = canonical.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=0, pool=0, threshold=0.354, k=k)
    a = canonical.assign(out["row"], out["score"], out["trust"], out["count"], 0.3, 2)
    parity = {"ids_equal": bool(np.array_equal(out["row"], ref[0])), "counts_equal": bool(np.array_equal(out["count"], ref[2])),
              "scores_bitequal": bool(np.array_equal(out["score"].view(np.uint32), ref[1].view(np.uint32))),
              "assignment_equal": bool(np.array_equal(out["assign_idx"], a[0]) and np.array_equal(out["assign_score"], a[1])),
              "oracle": "oracle/canonical.c (identify + the combine_signals restatement)"}
    gms, gl = ctx.profile_get("exact")
    pk, pk_src = peaks()
    avg_ms = gms / max(1, gl)
    roof = {"kernel": "k_exact_dense (fp64 SIMT, canonical arithmetic)", "bound": "hbm", "achieved": (N * D + P * D) * 4 / (avg_ms * 1e-3) / 1e9,
            "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": (N * D + P * D) * 4 / (avg_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
            "avg_launch_ms": avg_ms, "note": "33 KB of data: launch-latency bound, the fraction is informational (SURVEY 8d)"}
    ctx.close()
    # ---- through the real CLI, one process per step ----
    with tempfile.TemporaryDirectory(prefix="cfg1-") as td:
        audio, tpath, ids = synth.write_store(case, td, ["S1", "S2"], speaker_names=["alice", "bob", "carol"])
        env = dict(os.environ, SPEAKERS_EMBEDDINGS_DIR=td, SPEAKER_DETECTION_BACKEND="b200", PYTHONPATH=str(ROOT))
        env.pop("SPEAKER_BACKENDS_CONFIG", None)
        cmd = [sys.executable, str(ROOT / "bin" / "speaker-assign"), "-q", "assign", str(audio), "-t", str(tpath), "-e", "-n", "--format", "json",
               "--threshold", "0.2"]
        subprocess.run(cmd, env=env, capture_output=True, text=True)                      # warm-up (page cache, cubin load)
        e_steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        t_cli = (time.perf_counter() - t0) / e_steps
        got = json.loads(r.stdout[r.stdout.index("{"):]) if r.returncode == 0 else {}
        mapping = {l: m.get("speaker_id") for l, m in got.get("mappings", {}).items()}
        trust = ["high", "medium", "low"]
        want = {}
        for g, label in enumerate(["S1", "S2"]):
            sigs = [mnp.Signal("embedding_match", ids[case.row_speaker[int(ref[0][g, i])]], float(ref[1][g, i]),
                               {"trust_level": trust[case.row_trust[int(ref[0][g, i])]]}) for i in range(ref[2][g])]
            want[label] = mnp.combine_signals(label, sigs, threshold=0.2)["speaker_id"]
        parity["cli_mapping"] = mapping
        parity["cli_mapping_equal"] = mapping == want
        # time of an empty interpreter + package import, for scale
        t0 = time.perf_counter()
        subprocess.run([sys.executable, "-c", "import numpy"], env=env, capture_output=True)
        t_py = time.perf_counter() - t0
    parity["status"] = "ok" if all(v for k_, v in parity.items() if k_.endswith("equal")) else "MISMATCH"
    cpu = None
    if not args.no_cpu:
        goff = case.goff
        t0 = time.perf_counter()
        reps = 200
        for _ in range(reps):
            rows, scores, cnt = mnp.identify(case.seg, goff, case.bank, case.row_speaker, mode=0, pool=0, threshold=0.354, k=k)
            for g in range(L):
                mnp.combine_signals(f"S{g}", [mnp.Signal("embedding_match", str(int(rows[g, i])), float(scores[g, i]), {"trust_level": "high"})
                                              for i in range(cnt[g])], threshold=0.3)
        tc = (time.perf_counter() - t0) / reps
        cpu = {"value": pairs / tc, "unit": UNIT, "cores": blas_threads(), "kind": "port",
               "sample": f"the bare NumPy call, {reps} repetitions ({tc * 1e6:.0f} us each); the reference's CLI chain adds one process per label on top"}
    print(json.dumps({"metric": METRIC, "value": pairs * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(3, args.warmup),
                      "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic",
                      "config": {"workload": "configs[0]: speaker-assign on one synthetic Speechmatics-format transcript (2 labels, 40 segments) vs 3 enrolled profiles, 192-d",
                                 "segments": N, "labels": L, "bank_rows": P, "dim": D, "k": k, "l2": "256 MB flush buffer written between timed iterations",
                                 "path": "exact-simt"},
                      "clocks": clk, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
                      "e2e": {"value": pairs / t_cli, "unit": UNIT, "h2d_bytes_per_step": int(case.seg.nbytes + case.seg_label.nbytes // 2 + case.bank.nbytes),
                              "d2h_bytes_per_step": int(L * (k * 13 + 4)), "steps": e_steps, "seconds_per_cli_run": t_cli,
                              "python_numpy_import_s": t_py,
                              "sample": "one `bin/speaker-assign assign -e` PROCESS per step: interpreter start + CUDA context + store/sidecar reads + plugin + kernels + JSON"},
                      "parity_sample": parity, "time_to_solution_ms": ms / steps}))
    return 0


# ---- sampled full-scale parity (outside the timed region): the C oracle on randomly drawn label groups ----------------
def parity_sample(torch, dist, ctx, cfg, seg, lab, counts_flat, bank_local, world, rank, sharded, dev, n_groups=32, fma_budget=1.5e12,
                  as_f16=False):
    """Draws label groups of the batch this rank just scored, runs oracle/canonical.c (exhaustive canonical arithmetic,
    whole GLOBAL bank) on them and compares with what the GPU returned: rows and order, counts, scores bit for bit.
    Sharded bank: the shards are gathered to rank 0 first (the result is global).  Rank 0 reports."""
    P_local, D = bank_local.shape
    if sharded and world > 1:
        parts = [torch.empty_like(bank_local) for _ in range(world)] if rank == 0 else None
        dist.gather(bank_local, parts, dst=0)
        bank = torch.cat(parts) if rank == 0 else None
    else:
        bank = bank_local
    if rank != 0:
        return None
    from oracle import canonical
    out = ctx.fetch()
    goff = np.r_[0, np.cumsum(counts_flat)]
    G = len(counts_flat)
    P = bank.shape[0]
    rng = np.random.default_rng(cfg["seed"] + 77)
    order = rng.permutation(G)
    pick, fma = [], 0.0
    for g in order:                                    # random groups until the oracle's time budget is spent
        c = float(counts_flat[g]) * P * D
        if pick and fma + c > fma_budget:
            break
        pick.append(int(g))
        fma += c
        if len(pick) == n_groups:
            break
    pick.sort()
    h_bank = bank.cpu().numpy()
    if as_f16:      # what the fp16 host path scored: the embeddings rounded to fp16, widened back exactly
        segs = [seg[int(goff[g]):int(goff[g + 1])].to(torch.float16).float().cpu().numpy() for g in pick]
    else:
        segs = [seg[int(goff[g]):int(goff[g + 1])].cpu().numpy() for g in pick]
    sgoff = np.r_[0, np.cumsum([len(x) for x in segs])].astype(np.int64)
    t0 = time.perf_counter()
    ref = canonical.identify(np.concatenate(segs), sgoff, h_bank, np.arange(P, dtype=np.int32), P, mode=cfg["dtype"], pool=cfg.get("pool", 0),
                             threshold=cfg["thr"], k=cfg["k"])
    t_or = time.perf_counter() - t0
    ids_equal = bool(np.array_equal(out["row"][pick], ref[0]))
    return {"groups": len(pick), "segments": int(sgoff[-1]), "bank_rows": int(P), "ids_equal": ids_equal,
            "counts_equal": bool(np.array_equal(out["count"][pick], ref[2])),
            "scores_bitequal": bool(np.array_equal(out["score"][pick].view(np.uint32), ref[1].view(np.uint32))),
            "matched_groups": int((ref[2] > 0).sum()), "oracle": "oracle/canonical.c, exhaustive over the whole bank",
            "oracle_s": round(t_or, 2), "status": "ok" if ids_equal else "MISMATCH"}


# ---- identify workloads (cfg2, cfg3, cfg4*) ----------------------------------------------------------------------------
def run_identify(args, name, torch, dist, world, rank, local_rank, sub=False, stage_a=None, do_e2e=True):
    """One workload on `world` ranks: data, warm-up, timed steps, roofline, e2e, CPU baseline, sampled parity.
    Returns the JSON line as a dict on rank 0 (None elsewhere)."""
    from speaker_diarization_toolkit_b200 import _native
    cfg = dict(WORKLOADS[name])
    cfg["pool"] = 1 if args.pool == "max" else 0
    sharded = name.startswith("cfg4")
    cfg["scaling"] = "weak" if sharded else "strong"
    cfg["P_total"] = cfg["P"] * (world if sharded else 1)
    cfg["P_cpu"] = cfg["P_total"]
    counts, truth = group_counts(cfg, args.scale)
    if sharded:
        truth = np.where(truth >= 0, truth * world, truth)        # true speakers spread over all shards
    dev = torch.device("cuda", local_rank)
    uid = None
    if sharded and world > 1:
        box = [_native.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    ctx = _native.Context(local_rank, world if sharded else 1, rank if sharded else 0, uid)
    ctx.set_option("profile", 1)
    ctx.set_option("cta_group", args.cta_group)
    ctx.set_option("acc", args.acc)
    ctx.set_option("kth", args.kth)
    stage_a = stage_a or args.stage_a
    ctx.set_option("poolfirst", 1 if stage_a == "poolfirst" else 0)

    # ---- data: bank (replicated, or this rank's row shard) + this rank's recordings ----
    R = counts.shape[0]
    if sharded:
        r0, r1 = 0, R                                    # queries replicated, bank row-sharded
        cent, bank = make_bank(torch, cfg, dev, cfg["P"], row0=rank * cfg["P"])
        row_off = rank * cfg["P"]
    else:
        r0, r1 = rank * R // world, (rank + 1) * R // world   # recordings split over ranks
        cent, bank = make_bank(torch, cfg, dev, cfg["P"])
        row_off = 0
    P, D = bank.shape
    spk = torch.arange(P, device=dev, dtype=torch.int32) + (row_off if sharded else 0)      # one speaker per row, global ids
    trust = ((torch.arange(P, device=dev) + row_off) % 3).to(torch.uint8)
    ctx.bank_load_dev(bank.data_ptr(), spk.data_ptr(), trust.data_ptr(), P, D, cfg["dtype"], row_off)
    seg, lab, N, G = make_segments(torch, cfg, dev, cent, counts[r0:r1], truth[r0:r1], cfg["seed"] + (0 if sharded else 1000 + rank))
    del cent
    torch.cuda.synchronize()
    pairs_rank = float(N) * float(P)
    if sharded:
        pairs_total = float(N) * float(P) * world          # every rank scores all queries against its shard
    else:
        t = torch.tensor([pairs_rank], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        pairs_total = float(t.item())

    def step_dev():
        ctx.identify_dev(seg.data_ptr(), lab.data_ptr(), N, G, cfg["pool"], cfg["thr"], cfg["k"])
        ctx.assign(0.3, "low")

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step_dev()
    barrier()
    path, nfb = ctx.last_path()
    nretry = ctx.last_retry()
    ctx.profile_reset()
    l0 = ctx.launch_count()
    clocks.mark_begin()
    # working set of one step: raw + operand copies of segments and bank.  Below 2x L2 the inputs would stay cached
    # between steps, so every step is timed on its own after a 256 MB flush write
    Dp_ = (D + 63) // 64 * 64
    work_bytes = N * D * 4 + N * Dp_ * 2 + P * Dp_ * 2
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8) if work_bytes < (256 << 20) else None
    flush_rd = torch.zeros(256 << 18, device=dev, dtype=torch.int32) if flush is not None else None

    def l2_flush():
        """256 MB written (the contract's flush), then another 256 MB READ: the written lines would otherwise sit in L2 as
        dirty lines and their write-back would fall into the timed step (a plain 128 MB read pass measures 44 us right after
        the write and 21 us from a clean cache: profiles/r02_launches_cfg4i_i.csv)"""
        flush.zero_()
        flush_rd.sum()
        torch.cuda.synchronize()
    if flush is None:
        ctx.timer_start()
        for _ in range(args.steps):
            step_dev()
        ms = ctx.timer_stop()
    else:
        # latency-bound shapes: the step is timed WITHOUT the per-kernel event pairs (two events per launch would be a
        # third of a 50 us step); the kernel breakdown comes from a second, untimed round of the same steps
        ctx.set_option("profile", 0)
        ms = 0.0
        for _ in range(args.steps):
            l2_flush()
            ctx.timer_start()
            step_dev()
            ms += ctx.timer_stop()
        ctx.set_option("profile", 1)
        ctx.profile_reset()
        for _ in range(args.steps):
            l2_flush()
            step_dev()
        ctx.sync()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launch_count() - l0
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = pairs_total * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel, timed live with CUDA events on the library stream ----
    pk, pk_src = peaks()
    names = ["plan", "normalize", "poolgemm", "merge", "exact", "select", "assign", "allgather", "merge_topk"]
    prof = {n: ctx.profile_get(n) for n in names}
    tot_prof = sum(v[0] for n, v in prof.items() if n != "merge_topk") or 1.0     # merge_topk is nested inside allgather
    if path == 4:
        # bank-stream stage A (gemv.cu): one pass over the bf16 bank, HBM-bound; algorithmic bytes = 2*Dp per bank row
        gms, gl = prof["poolgemm"]
        avg_ms = gms / max(1, gl)
        Dp4 = (D + 63) // 64 * 64
        ach = (P * Dp4 * 2 + N * Dp4 * 2) / (avg_ms * 1e-3) / 1e9
        # the practical ceiling: one plain 128-bit read pass over the same bank on this GPU, same flush, same timer
        pms = 0.0
        for _ in range(5):
            l2_flush() if flush is not None else torch.cuda.synchronize()
            ctx.timer_start()
            ctx.probe_bank_read()
            pms += ctx.timer_stop()
        pms /= 5
        roof = {"kernel": "k_gemv8 (label offsets + normalise + bank stream through mma.sync m16n8k16 + per-CTA top lists, <= 8 query segments)", "bound": "hbm", "achieved": ach,
                "read_probe": {"ms": pms, "gbs": P * Dp4 * 2 / (pms * 1e-3) / 1e9,
                               "what": "k_probe_read: plain ld.global.cs.v4 pass over the same bank operands, timed the same way (events around one launch)"},
                "frac_of_read_probe": pms / avg_ms,
                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                "peak_source": f"{pk_src} copy bandwidth", "algorithmic_bytes_per_bank_row": 2 * Dp4, "avg_launch_ms": avg_ms,
                "share_of_step": gms / tot_prof}
    elif path == 5:
        # pool-first: a different algorithm (stage A contracts G centroids, not N segments): the step is HBM-bound on K1,
        # which reads every raw segment once and writes its bf16 operand copy; algorithmic bytes = 4*D + 2*Dp per segment
        gms, gl = prof["normalize"]
        avg_ms = gms / max(1, gl // 1)
        Dp5 = (D + 63) // 64 * 64
        per_launch = N * (4 * D + 2 * Dp5) / max(1, gl // max(1, args.steps))
        ach = per_launch / (avg_ms * 1e-3) / 1e9
        roof = {"kernel": "k_normalize_centroid (K1 + label centroids; stage A = centroid GEMM)", "bound": "hbm", "achieved": ach,
                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": f"{pk_src} copy bandwidth",
                "algorithmic_bytes_per_segment": 4 * D + 2 * Dp5, "avg_launch_ms": avg_ms, "share_of_step": gms / tot_prof,
                "whole_step_vs_hbm_floor": (N * 4 * D / (pk["hbm_gbs"] * 1e9)) / (1e-3 * ms_max / args.steps),
                "contracted_pairs_per_step": float(2 * G) * P,
                "note": "pairs/s below counts the segment x profile pairs of the WORKLOAD (what the step answers for), not pairs contracted"}
    elif path >= 2:
        gms, gl = prof["poolgemm"]
        per_launch_flops = 2.0 * pairs_rank * D / max(1, gl // max(1, args.steps))   # flops per launch = 2*D per pair
        avg_ms = gms / max(1, gl)
        ach = per_launch_flops / (avg_ms * 1e-3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        kname = f"k_poolacc{'' if args.cta_group == 1 else '2'} (tcgen05 cta_group::{1 if args.cta_group == 1 else 2}, mean pooling inside the MMA accumulation)" if path == 3 else \
            f"k_poolgemm (tcgen05 cta_group::{2 if args.cta_group == 2 else 1}, {args.pool} pooling in the epilogue)"
        roof = {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None, "peak_source": f"{pk_src} bf16 sustained (kernel timed inside a long step)",
                "frac_of_burst_peak": ach / pk["bf16_tflops"], "algorithmic_flop_per_pair": 2 * D,
                "avg_launch_ms": avg_ms, "share_of_step": gms / tot_prof}
    else:
        gms, gl = prof["exact"]
        avg_ms = gms / max(1, gl)
        bytes_alg = (N * D + P * D) * (2 if cfg["dtype"] else 4)
        ach = bytes_alg / (avg_ms * 1e-3) / 1e9
        roof = {"kernel": "k_exact_q30 (fp64 SIMT)", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": f"{pk_src} copy bandwidth",
                "avg_launch_ms": avg_ms, "share_of_step": gms / tot_prof,
                "note": "latency-bound shape: the fraction is informational (SURVEY 8d)"}
    roof["traffic"] = ncu_traffic(name + ("_max" if cfg["pool"] else "") + ("_poolfirst" if path == 5 else ""), roof["kernel"],
                                  args.scale == 1.0 and (world == 1 or sharded))
    if roof["traffic"] is not None:
        roof["traffic_source"] = "profiles/roofline_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this workload (not re-measured in this run)"

    # ---- certificate: measured stage-A error of the sampled candidates is reported by the parity tests; here: the margin model ----
    cert = None
    if path >= 2:
        try:
            _, _, eb, ec = ctx.stage_a()
            cert = {"eps_base": eb, "eps_per_chain_unit": ec, "fallback_groups": nfb, "retry_groups": nretry}
        except Exception:
            cert = None

    # ---- sampled parity at full scale (outside every timed region) ----
    par = None
    if not args.no_parity:
        step_dev()
        barrier()
        par = parity_sample(torch, dist, ctx, cfg, seg, lab, counts[r0:r1].reshape(-1), bank, world, rank, sharded, dev)
        barrier()

    # ---- e2e: host buffers through the host-pointer C-ABI call ----
    # Headline e2e: the segment embeddings sit in pinned host memory in the sidecar's compact fp16 form
    # (store.save_segment_embeddings(dtype=float16); sdk_identify_f16 widens them exactly on the device), so a step moves
    # 2*D bytes per segment over PCIe; `e2e_f32` is the same call on fp32 host embeddings (4*D bytes per segment).
    # Each rank's pinned buffers are allocated while the process is bound to the CPUs of ITS GPU's NUMA node.
    e2e, e2e_f32, par_e2e = None, None, None
    if not args.no_e2e and do_e2e:
        import psutil
        numa, numa_how = bind_to_gpu_numa_node(torch, local_rank)

        def e2e_leg(f16):
            esz = 2 if f16 else 4
            need = N * D * esz
            frac = 1.0
            avail = psutil.virtual_memory().available
            if need * 3 > avail:
                frac = max(0.05, avail / (need * 3.0))
            n_e = int(N * frac)
            if frac < 1.0:                                     # cut at a label-group boundary
                n_e = int((lab[:n_e] != lab[n_e - 1]).sum().item())
            g_e = int(lab[n_e - 1].item()) + 1 if n_e else 0
            h_seg = torch.empty((n_e, D), dtype=torch.float16 if f16 else torch.float32, pin_memory=True)
            h_lab = torch.empty((n_e,), dtype=torch.int32, pin_memory=True)
            step_rows = 1 << 21
            for a in range(0, n_e, step_rows):                 # converted slice by slice: no second full-size device copy
                b = min(n_e, a + step_rows)
                h_seg[a:b].copy_(seg[a:b].to(h_seg.dtype) if f16 else seg[a:b])
            h_lab.copy_(lab[:n_e])
            torch.cuda.synchronize()
            hs, hl = h_seg.numpy(), h_lab.numpy()

            def step_host():
                ctx.identify(hs, hl, g_e, pool=cfg["pool"], threshold=cfg["thr"], k=cfg["k"])
                ctx.assign(0.3, "low")
                return ctx.fetch(with_assign=True)

            step_host()
            barrier()
            e_steps = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            for _ in range(e_steps):
                step_host()
            ctx.sync()
            te = time.perf_counter() - t0
            tt = torch.tensor([te], device=dev, dtype=torch.float64)
            pe = torch.tensor([float(n_e) * float(P) * (world if sharded else 1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                if not sharded:
                    dist.all_reduce(pe)
            k = cfg["k"]
            rec = {"value": float(pe.item()) * e_steps / float(tt.item()), "unit": UNIT,
                   "h2d_bytes_per_step": int(n_e * D * esz + n_e * 4), "d2h_bytes_per_step": int(g_e * (k * 13 + 4 + 4 + 8 + 4 + 12 + 24)),
                   "steps": e_steps, "sample": "whole batch" if frac == 1.0 else f"first {frac:.2f} of the rank's batch (host memory bound)",
                   "input": ("fp16 segment embeddings in pinned host memory (compact sidecar form, widened exactly on the device)" if f16
                             else "fp32 segment embeddings in pinned host memory"),
                   "h2d_gbs_per_rank": (n_e * D * esz + n_e * 4) * e_steps / float(tt.item()) / 1e9, "numa_node_of_gpu": numa,
                   "numa_placement": numa_how}
            par = None
            if f16 and not args.no_parity:
                # the results of the last fp16 host step against the oracle on the widened fp16 values
                par = parity_sample(torch, dist, ctx, cfg, seg[:n_e], lab[:n_e], counts[r0:r1].reshape(-1)[:g_e], bank, world, rank, sharded, dev,
                                    n_groups=16, as_f16=True)
            del h_seg, h_lab, hs, hl
            return rec, par

        if args.e2e_dtype in ("f16", "both"):
            e2e, par_e2e = e2e_leg(True)
        if args.e2e_dtype in ("f32", "both"):
            e2e_f32, _ = e2e_leg(False)
            if e2e is None:
                e2e, e2e_f32 = e2e_f32, None
        unbind_cpus()

    # ---- CPU baseline (rank 0, bounded sample) ----
    cpu = None
    if rank == 0 and not args.no_cpu and world == 1 and not sub:
        rate, n_rec, t = cpu_port_rate(cfg, counts, truth, 12.0)
        cpu = {"value": rate, "unit": UNIT, "cores": blas_threads(), "host_cores": os.cpu_count(), "kind": "port",
               "sample": f"{n_rec} recordings of the same workload in {t:.1f} s (NumPy/OpenBLAS sgemm + combine_signals), extrapolates linearly"}

    line = None
    if rank == 0:
        ag_ms, ag_n = prof["allgather"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
                "dtype": "bf16" if cfg["dtype"] else "f32", "data": "synthetic",
                "config": {"workload": cfg["desc"], "segments_total": int(pairs_total / (P * (world if sharded else 1))),
                           "segments_per_gpu": N, "label_groups_per_gpu": G, "bank_rows_per_gpu": P, "dim": D, "k": cfg["k"],
                           "threshold": cfg["thr"], "pool": args.pool, "stage_a": stage_a, "parallelism": ("bank-row-sharded x" if sharded else "dp") + str(world),
                           "l2": "working set larger than 2x L2 (no flush needed)" if flush is None else "256 MB flush buffer written, then 256 MB read (no dirty lines left), between timed iterations",
                           "path": {1: "exact-simt", 2: f"tcgen05 cta_group::{2 if args.cta_group == 2 else 1}", 3: f"tcgen05 accumulate-pooling cta_group::{1 if args.cta_group == 1 else 2}", 4: "bank-stream gemv", 5: "pool-first (label centroids x bank on tcgen05, then the canonical re-score)"}.get(path, str(path)), "certificate_fallback_groups": nfb, "certificate_retry_groups": nretry, "scale": args.scale},
                "clocks": clk, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_f32": e2e_f32,
                "parity_sample": par, "parity_sample_e2e": par_e2e, "certificate": cert,
                "time_to_solution_ms": ms_max / args.steps,
                "kernel_ms_per_step": {n: v[0] / args.steps for n, v in prof.items()}}
        if sharded and world > 1:
            line["allgather_merge_us"] = 1e3 * ag_ms / max(1, ag_n)
    ctx.close()
    del seg, lab, bank
    torch.cuda.empty_cache()
    return line


def sharded_subrecord(sub):
    """What the top-level line keeps of the row-sharded (configs[3]) run that follows the default workload."""
    r = sub["roofline"]
    return {"workload": sub["config"]["workload"], "value": sub["value"], "unit": UNIT, "ms_per_step": sub["ms_per_step"],
            "n_gpus": sub["n_gpus"], "scaling": "weak", "bank_rows_total": sub["config"]["bank_rows_per_gpu"] * sub["n_gpus"],
            "segments": sub["config"]["segments_per_gpu"], "label_groups": sub["config"]["label_groups_per_gpu"],
            "roofline": {"kernel": r["kernel"], "bound": r["bound"], "achieved": r["achieved"], "peak": r["peak"], "unit": r["unit"],
                         "frac": r["frac"], "avg_launch_ms": r["avg_launch_ms"], "share_of_step": r["share_of_step"], "traffic": r["traffic"],
                         "whole_step_frac": 2.0 * sub["config"]["dim"] * sub["value"] / sub["n_gpus"] / 1e12 / r["peak"] if r["bound"] == "tensor" else None},
            "allgather_merge_us": sub.get("allgather_merge_us"), "parity_sample": (sub["parity_sample"] or {}).get("status"),
            "parity": sub["parity_sample"], "e2e": sub["e2e"], "e2e_f32": sub.get("e2e_f32"), "clocks": sub["clocks"], "gpu_launches": sub["gpu_launches"],
            "certificate": sub["certificate"], "path": sub["config"]["path"], "kernel_ms_per_step": sub["kernel_ms_per_step"]}


# ---- main -----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--pool", default="mean", choices=["mean", "max"], help="per-label pooling (north_star (3): mean/max)")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the recordings (debug runs only; 1.0 = the named config)")
    ap.add_argument("--cta-group", type=int, default=0, choices=[0, 1, 2], help="tcgen05 kernel variant: 1 single CTA, 2 CTA pairs, 0 auto (pairs for accumulate-pooling)")
    ap.add_argument("--acc", type=int, default=1, choices=[0, 1, 2], help="accumulate-pooling kernel: 0 off, 1 auto, 2 force (A/B runs)")
    ap.add_argument("--kth", type=int, default=1, choices=[0, 1, 2], help="running k-th best pruning of the candidate flush: 0 off, 1 auto, 2 on (A/B runs)")
    ap.add_argument("--stage-a", default="contraction", choices=["contraction", "poolfirst"],
                    help="poolfirst: stage A contracts the label centroids (mean pooling only; a different algorithm, HBM-bound)")
    ap.add_argument("--no-poolfirst", action="store_true", help="do not append the pool-first sub-record to the default workload's line")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-dtype", default="both", choices=["f16", "f32", "both"],
                    help="storage type of the host segment embeddings in the e2e leg (f16 = compact sidecar form, the headline; both = also e2e_f32)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled oracle check that follows the timed region")
    ap.add_argument("--no-sharded", action="store_true", help="do not append the row-sharded configs[3] run to the default workload's line")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        cfg = dict(WORKLOADS[args.workload])
        cfg["pool"] = 1 if args.pool == "max" else 0
        sharded = args.workload.startswith("cfg4")
        cfg["scaling"] = "weak" if sharded else "strong"
        cfg["P_total"] = cfg["P"] * (world if sharded else 1)
        cfg["P_cpu"] = cfg["P_total"]
        counts, truth = group_counts(cfg, args.scale)
        if sharded:
            truth = np.where(truth >= 0, truth * world, truth)
        return run_reference(args, cfg, counts, truth)
    if args.workload == "cfg5":
        return run_cfg5(args, dict(WORKLOADS["cfg5"]))
    if args.workload == "cfg1":
        return run_cfg1(args)

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = run_identify(args, args.workload, torch, dist, world, rank, local_rank)
    # The default workload (configs[2], data-parallel, no collective) is followed by the row-sharded one (configs[3]:
    # 125k bank rows per GPU, one ncclAllGather + merge per step) so that every driver run of `bench.py --gpus N` also
    # measures the path north_star sets its target on.
    if args.workload == "cfg3" and not args.no_sharded and args.scale == 1.0:
        sub = run_identify(args, "cfg4", torch, dist, world, rank, local_rank, sub=True)
        if rank == 0:
            line["sharded"] = sharded_subrecord(sub)
    # ... and, for mean pooling, by the same workload with the pool-first stage A: a DIFFERENT algorithm, reported
    # beside the contraction as time to solution against the HBM roofline -- never as the headline pairs/s
    if args.workload == "cfg3" and args.pool == "mean" and args.stage_a == "contraction" and not args.no_poolfirst:
        pf = run_identify(args, "cfg3", torch, dist, world, rank, local_rank, sub=True, stage_a="poolfirst", do_e2e=False)
        if rank == 0:
            line["pool_first"] = {"algorithm": "stage A contracts the label centroids (mean pooling is linear); stage B re-scores the candidates over all segments in the canonical arithmetic",
                                  "time_to_solution_ms": pf["ms_per_step"], "contraction_time_to_solution_ms": line["ms_per_step"],
                                  "workload_pairs_per_s": pf["value"], "roofline": pf["roofline"], "parity_sample": pf["parity_sample"],
                                  "certificate": pf["certificate"], "kernel_ms_per_step": pf["kernel_ms_per_step"], "clocks": pf["clocks"],
                                  "gpu_launches": pf["gpu_launches"], "path": pf["config"]["path"]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
