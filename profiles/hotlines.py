#!/usr/bin/env python
"""Per-source-line instruction and stall-sample table of one kernel of an .ncu-rep (`ncu --set full --import-source on`).

    python profiles/hotlines.py REPORT.ncu-rep MANGLED_NAME_PATTERN [top N] [--by exec|samples|hot]

The report's per-SASS-instruction counters (`--page source --print-source sass`) are joined with the line table of the
library as built NOW (`cuobjdump -xelf` + `nvdisasm -c -g` of python-audio-tools_b200/libb200flac.so): instruction i of
the function in the report is instruction i of the listing.  Only meaningful while the library is the one profiled
(the script checks that the instruction counts agree).  exec = executions per warp and CTA; hot = SASS instructions
executed at least 0.2 times per warp and CTA (the instruction-cache working set, see r02_icache_analysis.txt).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("B200FLAC_LIB") or os.path.join(ROOT, "python-audio-tools_b200", "libb200flac.so")
SRC = os.path.join(ROOT, "python-audio-tools_b200", "csrc")


def listing(pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    ins, name = [], None
    for cubin in sorted(os.listdir(tmp)):
        text = subprocess.run(["nvdisasm", "-c", "-g", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
        on, cur = False, None
        for l in text.split("\n"):
            m = re.match(r"\.text\.(\S+):", l)
            if m:
                if on:
                    break
                on = re.search(pattern, m.group(1)) is not None
                if on:
                    name = m.group(1)
                continue
            if not on:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
            if m:
                ins.append((m.group(2).strip(), cur))
        if ins:
            break
    return name, ins


def counters(rep, demangled_hint):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    blocks, cur, hdr, name = [], None, None, None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append((r[1], cur))
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None and len(r) > 6 and r[0].startswith("0x"):
            cur.append(r)
    for name, rows in blocks:
        if re.search(demangled_hint, name):
            return name, hdr, rows
    raise SystemExit("no kernel matching %r in %s" % (demangled_hint, rep))


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
    by = sys.argv[sys.argv.index("--by") + 1] if "--by" in sys.argv else "samples"
    mangled, ins = listing(pattern)
    if not ins:
        raise SystemExit("no function matching %r in %s" % (pattern, LIB))
    base_name = re.match(r"_Z\d+([A-Za-z_0-9]+?)I", mangled).group(1)
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout)))
    h = raw[0]
    row = [r for r in raw[2:] if base_name in r[h.index("Kernel Name")]][0]
    name, hdr, rows = counters(rep, base_name)
    if len(rows) != len(ins):
        raise SystemExit("the library has %d instructions in %s, the report %d: not the build that was profiled" % (len(ins), mangled, len(rows)))
    base = float(row[h.index("launch__grid_size")].replace(",", "")) * float(row[h.index("launch__block_size")].replace(",", "")) / 32.0
    ix, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    ex, sm, n, hot, ops = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for (text, line), r in zip(ins, rows):
        e = int(r[ix])
        ex[line] += e
        sm[line] += int(r[ismp])
        n[line] += 1
        if e >= 0.2 * base:
            hot[line] += 1
        t = text.split()
        ops[line][t[1] if t[0].startswith("@") else t[0]] += e
    tote, tots = float(sum(ex.values())), float(sum(sm.values()))
    src = {}
    print("%s\n%s: %d SASS instructions, %d hot; %.0f executed per warp and CTA; %d stall samples"
          % (name[:140], os.path.basename(rep), len(ins), sum(hot.values()), tote / base, tots))
    key = {"exec": ex, "samples": sm, "hot": hot}[by]
    for line, _ in key.most_common(topn):
        text = ""
        if line:
            path = os.path.join(SRC, line[0])
            if line[0] not in src:
                src[line[0]] = open(path, errors="replace").read().split("\n") if os.path.exists(path) else []
            if line[1] <= len(src[line[0]]):
                text = src[line[0]][line[1] - 1].strip()[:70]
        top = " ".join("%s:%.0f" % (o, c / base) for o, c in ops[line].most_common(3))
        print("%-18s %4d  sass %4d hot %4d  exec %8.1f (%4.1f%%)  samples %4.1f%%  | %-70s | %s"
              % (line[0][:18] if line else "-", line[1] if line else 0, n[line], hot[line], ex[line] / base,
                 100 * ex[line] / tote, 100 * sm[line] / tots, text, top))


if __name__ == "__main__":
    main()
