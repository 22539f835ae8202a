"""Per-launch table of an ncu report: python tools/ncu_table.py report.ncu-rep [out.md]"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("launch__grid_size", "grid", 1), ("launch__block_size", "blk", 1),
        ("launch__registers_per_thread", "regs", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("lts__t_sector_hit_rate.pct", "l2hit%", 1), ("dram__bytes_read.sum", "rdMB", 1e-6), ("dram__bytes_write.sum", "wrMB", 1e-6),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tens%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1)]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
lines = ["| # | kernel | " + " | ".join(c[1] for c in COLS) + " |", "|---|---|" + "---|" * len(COLS)]
for k, r in enumerate(rows[2:]):
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("aesr::", "")
    vals = []
    for key, _, scale in COLS:
        try:
            v = float(r[idx[key]].replace(",", "")) * scale
            u = units[idx[key]]
            if key == "gpu__time_duration.sum":
                v = float(r[idx[key]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
            if key.startswith("dram__bytes"):
                v = float(r[idx[key]].replace(",", "")) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            vals.append("%.1f" % v)
        except Exception:
            vals.append("-")
    lines.append("| %d | `%s` | %s |" % (k, name[:60], " | ".join(vals)))
text = "\n".join(lines)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
else:
    print(text)
