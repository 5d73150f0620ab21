#!/usr/bin/env python
"""SASS evidence for the hot kernels: opcode histogram per kernel and the densest stretch of each
kernel's characteristic instruction.

    python profiles/sass_report.py python-audio-tools_b200/libb200flac.so > profiles/r01_v4_sass.txt
"""
import collections
import re
import subprocess
import sys

HOT = ["k_lpc_autocILi12ELi1E", "k_lpc_autocILi12ELi2E", "k_lpc_finishILi12E", "k_analyze_v3ILi5E", "k_pack_v3ILi256ELi4E",
       "k_frame_select", "k_scan_offsets"]
LOOPS = {"k_lpc_autocILi12ELi1E": "DFMA", "k_analyze_v3ILi5E": "IMAD", "k_pack_v3ILi256ELi4E": "ATOMS"}


def strip_pred(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t)


def main():
    so = sys.argv[1]
    txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((m.group(1), m.group(2).strip()))
    for name, ins in funcs.items():
        if not any(h in name for h in HOT):
            continue
        print("=" * 110)
        print(name)
        print("  %d instructions (%.0f KB)" % (len(ins), len(ins) * 16 / 1024.0))
        hist = collections.Counter()
        for _, t in ins:
            hist[strip_pred(t).split()[0].split(".")[0]] += 1
        print("  opcodes: " + "  ".join("%s %d" % kv for kv in hist.most_common(28)))
        tc = [k for k in hist if k.startswith(("HMMA", "IMMA", "UTC", "TCGEN", "QMMA", "DMMA"))]
        print("  tensor-core opcodes: %s" % (", ".join(tc) if tc else "none (integer/FP64 path, see DESIGN.md)"))
        key = next((v for k, v in LOOPS.items() if k in name), None)
        if key:
            flags = [1 if strip_pred(t).startswith(key) else 0 for _, t in ins]
            best, bi, run = -1, 0, sum(flags[:48])
            for i in range(len(ins) - 48):
                if run > best:
                    best, bi = run, i
                run += flags[i + 48] - flags[i]
            print("  densest %s stretch (48 instructions from /*%s*/):" % (key, ins[bi][0]))
            for a, t in ins[bi:bi + 48]:
                print("      /*%s*/  %s ;" % (a, t))


if __name__ == "__main__":
    main()
