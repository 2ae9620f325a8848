"""Compact per-launch summary of an `ncu --set full` report (run where ncu is installed, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1_global_ld_bytes"),
    ("lts__t_bytes.sum", "l2_bytes"),
]


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        parts = ["%s" % name[:90]]
        for k, label in KEYS:
            if k in col:
                parts.append("%s=%s%s" % (label, r[col[k]], (" " + units[col[k]]) if units[col[k]] else ""))
        lines.append("\n    ".join(parts))
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write("source: %s\n%s\n" % (rep, txt))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
