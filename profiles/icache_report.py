#!/usr/bin/env python
"""Instruction-fetch report of the kernels in an .ncu-rep captured with `ncu --set full --import-source on`:
SM instruction-cache hit rate, GPC instruction-cache (GCC) request rate, issue utilisation, the instruction-fetch
and barrier stalls, and the size of the HOT code -- SASS instructions executed at least 0.2 times per warp and
work item -- in instructions and in 128-byte instruction lines.

    python profiles/icache_report.py gpurun_out/prof_r02_final.ncu-rep [more.ncu-rep ...] > profiles/r02_icache_analysis.txt

Per-source-line numbers (which lines the hot instructions come from) are obtained by joining the same per-instruction
counts with `nvdisasm -c -g` of the cubin extracted from libb200flac.so (`cuobjdump -xelf all`): the i-th SASS
instruction of the function in the report is the i-th instruction of the listing, and `//## File ... line N` comments
give its source line.
"""
import csv
import re
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
           "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread"]


def norm(name):
    return re.sub(r"\((int|bool|unsigned int)\)", "", name)


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                             stdout=subprocess.PIPE, text=True).stdout
        kernels, cur, shdr = {}, None, None      # the page lists every kernel once per view: keep the first
        for r in csv.reader(io.StringIO(src)):
            if r and r[0] == "Kernel Name":
                cur = [] if norm(r[1]) not in kernels else None
                if cur is not None:
                    kernels[norm(r[1])] = cur
            elif r and r[0] == "Address":
                shdr = r
            elif cur is not None and len(r) > 6 and r[0].startswith("0x"):
                cur.append(r)
        print("=" * 110)
        print(rep.split("/")[-1])
        for k, r in enumerate(rows[2:]):
            print("-" * 110)
            print(r[hdr.index("Kernel Name")][:150])
            for m in METRICS:
                if m in hdr:
                    print("  %-82s %s" % (m, r[hdr.index(m)]))
            if norm(r[hdr.index("Kernel Name")]) in kernels and shdr:
                ix = shdr.index("Instructions Executed")
                grid = float(r[hdr.index("launch__grid_size")].replace(",", ""))
                warps = float(r[hdr.index("launch__block_size")].replace(",", "")) / 32.0
                base = grid * warps
                v = kernels[norm(r[hdr.index("Kernel Name")])]
                hot = [i for i, x in enumerate(v) if int(x[ix]) >= 0.2 * base]
                lines = set(i // 8 for i in hot)
                print("  SASS instructions %d, hot (>= 0.2 executions per warp and CTA) %d = %.1f KB, hot 128-byte lines %d = %.1f KB"
                      % (len(v), len(hot), len(hot) * 16 / 1024.0, len(lines), len(lines) * 128 / 1024.0))


if __name__ == "__main__":
    main()
