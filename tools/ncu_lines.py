#!/usr/bin/env python
"""Summarise an ncu report's source page per CUDA source line (needs -lineinfo + --import-source on).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--top 40] [--per N]

Prints, per source line, warp-instructions executed (optionally divided by --per, e.g. the number of sources the
launch processed), stall samples and the dominant stall reasons; then an opcode histogram of the whole kernel.
"""
import csv
import subprocess
import sys
from collections import Counter, defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    per = float(sys.argv[sys.argv.index("--per") + 1]) if "--per" in sys.argv else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr = None, None
    lines = {}          # (file, line) -> dict
    ops = Counter()
    cur = None
    tot_inst = tot_samp = 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            ix = {n: i for i, n in enumerate(hdr)}
            i_inst = ix["Instructions Executed"]; i_samp = ix["# Samples"]
            stall_cols = [(n, i) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0]:          # a CUDA source line: aggregated values follow
            cur = (fname, int(r[0]))
            lines[cur] = dict(src=r[1].strip(), inst=0, samp=0, stalls=Counter(), nsass=0)
            continue
        if r[2] in ("", "...") or cur is None:
            continue
        try:
            inst = int(r[i_inst]); samp = int(r[i_samp])
        except ValueError:
            continue
        d = lines[cur]
        d["inst"] += inst; d["samp"] += samp; d["nsass"] += 1
        for n, i in stall_cols:
            try:
                d["stalls"][n[6:]] += int(r[i])
            except ValueError:
                pass
        sass = r[3].strip()
        if sass.startswith("@"):
            sass = sass.split(None, 1)[1]
        ops[sass.split()[0].split(".")[0]] += inst
        tot_inst += inst; tot_samp += samp
    print("total warp-instructions %d (%.1f per unit), samples %d" % (tot_inst, tot_inst / per, tot_samp))
    print("\n== top lines by instructions executed ==")
    for (f, ln), d in sorted(lines.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        st = ", ".join("%s %d" % kv for kv in d["stalls"].most_common(3))
        print("%-14s %5d  inst %9.1f  samp %6d (%4.1f%%)  [%s]  %s" % (f, ln, d["inst"] / per, d["samp"], 100.0 * d["samp"] / max(1, tot_samp), st, d["src"][:90]))
    print("\n== top lines by stall samples ==")
    for (f, ln), d in sorted(lines.items(), key=lambda kv: -kv[1]["samp"])[:top]:
        st = ", ".join("%s %d" % kv for kv in d["stalls"].most_common(4))
        print("%-14s %5d  inst %9.1f  samp %6d (%4.1f%%)  [%s]  %s" % (f, ln, d["inst"] / per, d["samp"], 100.0 * d["samp"] / max(1, tot_samp), st, d["src"][:90]))
    print("\n== opcodes ==")
    for o, c in ops.most_common(30):
        print("  %-10s %10.1f" % (o, c / per))


if __name__ == "__main__":
    main()
