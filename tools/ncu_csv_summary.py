"""Compact per-launch summary of `ncu --page raw --csv` output: python tools/ncu_csv_summary.py in.csv out.txt"""
import csv, sys
KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "")[:70]
    parts = []
    for k, lab in KEYS:
        if k in col:
            v = r[col[k]]
            u = units[col[k]]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            parts.append("%s=%s%s" % (lab, v, (" " + u) if u and u not in ("%",) and lab not in ("regs", "grid", "block") else ""))
    out.append("%-70s %s" % (name, "  ".join(parts)))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("source: ncu --set full --clock-control none on `python bench.py --steps 2 --profile` (B200), raw page\n" + txt + "\n")
