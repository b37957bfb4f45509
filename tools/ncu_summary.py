#!/usr/bin/env python
"""Text summary of one `ncu --set full` capture for profiles/: headline metrics, stall-reason shares, opcode mix per
unit of work, and the source lines that cost the most (tools/ncu_lines.py gives the long form).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep --per <units in the launch> [--unit source] > profiles/rNN_*.txt
"""
import csv
import subprocess
import sys
from collections import Counter

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    per = float(sys.argv[sys.argv.index("--per") + 1]) if "--per" in sys.argv else 1.0
    unit = sys.argv[sys.argv.index("--unit") + 1] if "--unit" in sys.argv else "unit"
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    h = rows[0]
    ix = {n: i for i, n in enumerate(h)}
    v = rows[2]
    print("capture: %s   (ncu --set full --clock-control none; one launch; cold-cache, serialised replays -- shares, not absolutes)" % rep.split("/")[-1])
    print("kernel : %s" % v[ix["Kernel Name"]])
    vals = {}
    for n in WANT:
        if n in ix:
            vals[n] = v[ix[n]]
            print("  %-72s %-16s %s" % (n, rows[1][ix[n]], v[ix[n]]))

    def num(name):
        try:
            return float(vals[name])
        except Exception:
            return None
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    ru, wu = rows[1][ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_write.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if rd is not None and wr is not None:
        tb = rd * scale.get(ru, 1.0) + wr * scale.get(wu, 1.0)
        print("  traffic (dram read + write) per launch: %.0f bytes = %.1f bytes per %s" % (tb, tb / per, unit))
    inst = num("smsp__inst_executed.sum")
    if inst:
        print("  warp-instructions per %s: %.1f" % (unit, inst / per))
    src = list(csv.reader(ncu(rep, "--page", "source", "--csv").splitlines()))
    hdr = src[1]
    sx = {n: i for i, n in enumerate(hdr)}
    st, ops = Counter(), Counter()
    for r in src[2:]:
        for n, i in sx.items():
            if n.startswith("stall_") and "Not Issued" not in n:
                try:
                    st[n[6:]] += int(r[i])
                except ValueError:
                    pass
        try:
            n = int(r[sx["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        s = r[1].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        ops[s.split()[0].split(".")[0]] += n
    tot = sum(st.values()) or 1
    print("stall reasons (share of warp samples): " + ", ".join("%s %.1f%%" % (k, 100.0 * c / tot) for k, c in st.most_common(10)))
    print("opcode mix (warp-instructions per %s): " % unit + ", ".join("%s %.1f" % (k, c / per) for k, c in ops.most_common(22)))
    sass = " ".join(sorted(k for k in ops if k in ("UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "LDGSTS")))
    print("Blackwell / Hopper+ SASS present: " + sass)


if __name__ == "__main__":
    main()
