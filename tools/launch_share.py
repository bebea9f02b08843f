#!/usr/bin/env python3
"""Per-kernel totals and shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).

    python tools/launch_share.py profiles/r01_launches_v7.csv [--skip-first N]

ncu serialises launches and times them cold, so only each kernel's SHARE of the step is comparable with the CUDA-event
numbers of bench.py (`kernel_ms_per_step`)."""
import csv
import sys
from collections import defaultdict


def main() -> int:
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        rows.append((name, float(r["Metric Value"].replace(",", "")) * scale))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for n, ms in rows:
        tot[n] += ms
        cnt[n] += 1
    # the library's own kernels are named k_*; everything else in a bench.py capture is torch generating the synthetic data
    ours = {n: ms for n, ms in tot.items() if n.startswith("k_")}
    total = sum(ours.values())
    print(f"{len(rows)} launches; {sum(cnt[n] for n in ours)} of this library, {total:.3f} ms in them "
          f"(+ {sum(tot.values()) - total:.3f} ms of torch data-generation kernels, outside the timed region)")
    for n, ms in sorted(ours.items(), key=lambda kv: -kv[1]):
        print(f"{ms:10.3f} ms {100 * ms / total:6.2f} %  x{cnt[n]:<4d} {n}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
