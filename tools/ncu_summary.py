#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON that can be committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r01_poolacc_v6.ncu-rep profiles/r01_poolacc_v6_ncu_summary.json

Reads the raw page with `ncu -i REP --page raw --csv` (works without a GPU) and keeps, per profiled launch, the
metrics DESIGN.md / bench.py quote: duration, DRAM bytes, tensor-pipe and memory-pipe utilisation, launch shape.
With `--traffic-key WORKLOAD` the DRAM bytes of the longest profiled launch go into `profiles/roofline_traffic.json`,
which `bench.py` reads for the `roofline.traffic` field of that workload."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "sm__maximum_warps_per_active_cycle_pct",
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}


def main() -> int:
    if len(sys.argv) < 3:
        print(__doc__)
        return 2
    rep, out = sys.argv[1], sys.argv[2]
    traffic_key = None
    if "--traffic-key" in sys.argv:
        traffic_key = sys.argv[sys.argv.index("--traffic-key") + 1]
    traffic_path = "profiles/roofline_traffic.json"
    if "--traffic-out" in sys.argv:                 # e.g. gpurun_out/roofline_traffic.json when summarising on the GPU box
        traffic_path = sys.argv[sys.argv.index("--traffic-out") + 1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    launches = []
    for r in data:
        rec = {"kernel": r[col["Kernel Name"]], "id": r[col["ID"]]}
        for m in KEEP:
            if m in col and r[col[m]] != "":
                rec[m] = {"value": r[col[m]], "unit": units[col[m]]}
        launches.append(rec)
    summary = {"report": rep.split("/")[-1], "launches": launches}

    def scaled(rec, name):
        v = rec.get(name)
        if not v:
            return None
        return float(v["value"].replace(",", "")) * UNIT_SCALE.get(v["unit"], 1.0)

    for rec in launches:
        rd, wr = scaled(rec, "dram__bytes_read.sum"), scaled(rec, "dram__bytes_write.sum")
        if rd is not None and wr is not None:
            rec["dram_bytes_total"] = rd + wr
        ms = scaled(rec, "gpu__time_duration.sum")
        if ms is not None:
            rec["duration_ms"] = ms
    with open(out, "w") as fh:
        json.dump(summary, fh, indent=1)
    if traffic_key:
        path = traffic_path
        try:
            table = json.load(open(path))
        except (OSError, ValueError):
            table = {}
        # the longest profiled launch is the dominant kernel of the capture
        best = max((r for r in launches if "dram_bytes_total" in r), key=lambda r: r.get("duration_ms", 0.0), default=None)
        if best:
            table[traffic_key] = {"kernel": best["kernel"].split("(")[0].split("<")[0].replace("void ", "").strip(),
                                  "dram_bytes_per_launch": best["dram_bytes_total"],
                                  "duration_ms_under_ncu": best.get("duration_ms"), "report": summary["report"]}
            json.dump(table, open(path, "w"), indent=1)
    for rec in launches:
        print(rec["kernel"][:60], rec.get("duration_ms"), rec.get("dram_bytes_total"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
