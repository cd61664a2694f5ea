"""Summarise ncu outputs into profiles/ (tracked):
  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/launches_r01.md
  python tools/ncu_summary.py full gpurun_out/full_X_r01.ncu-rep profiles/full_X_r01.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("cm31::", "")


def launches(src, dst):
    rows = [r for r in csv.DictReader(l for l in open(src) if l.startswith('"'))]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"].replace(",", "")) / 1e6
    total = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list summary ({src})", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).", "",
           f"total {sum(v[0] for v in agg.values())} launches, {total:.3f} ms", "",
           "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f}% | {1e3 * ms / n:.1f} |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum",
           "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum", "launch__occupancy_limit_registers", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fmaheavy.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"]


def full(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [m for m in METRICS if m in idx]
    out = [f"# ncu --set full summary ({src})", "", "| # | kernel | grid | " + " | ".join(c.replace("__", " ").replace(".sum", "") for c in cols) + " |",
           "|---|---|---|" + "---|" * len(cols)]
    for r in rows:
        out.append(f"| {r[idx['ID']]} | {short(r[idx['Kernel Name']])} | {r[idx['Grid Size']]} | " + " | ".join(r[idx[c]] for c in cols) + " |")
    out.append("")
    out.append("units: " + ", ".join(f"{c}: {units[idx[c]]}" for c in cols))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
