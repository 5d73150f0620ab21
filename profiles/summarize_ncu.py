#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the per-kernel text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_r01.ncu-rep > profiles/r01_xxx_summary.txt
    python profiles/summarize_ncu.py gpurun_out/prof_r02.ncu-rep --json profiles/r02_traffic.json > ...

--json also writes, per kernel, the DRAM bytes of one launch (dram__bytes_read.sum + dram__bytes_write.sum),
its duration and issue-slot utilisation: the file bench.py reads `roofline.traffic` from.
"""
import json
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    def num(r, name):
        try:
            v = float(r[idx[name]])
        except (KeyError, ValueError):
            return None
        u = units[idx[name]].lower()
        scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
        return v * scale.get(u, 1.0)
    if "--json" in sys.argv:
        out = {"source": rep.split("/")[-1], "how": "ncu --set full --clock-control none, one launch per kernel", "kernels": {}}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            short = name.split("(")[0].replace("void ", "").strip()
            rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
            out["kernels"][short] = {"dram_bytes": (rd or 0) + (wr or 0), "dram_bytes_read": rd, "dram_bytes_write": wr,
                                     "duration_ms": num(r, "gpu__time_duration.sum"),
                                     "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                     "inst_executed": num(r, "smsp__inst_executed.sum")}
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
            json.dump(out, fh, indent=1)
    for r in rows[2:]:
        print("=" * 100)
        print(r[idx["Kernel Name"]][:160])
        for w in WANT:
            if w in idx:
                print("  %-68s %s %s" % (w, r[idx[w]], units[idx[w]]))
        top = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
        for v, h in top:
            print("  stall %-62s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    main()
