"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_conv_r01.ncu-rep profiles/r01_conv_full.md

`launches`: per-kernel totals and shares of a `--metrics gpu__time_duration.sum` launch list (cold-cache, serialised:
the SHARES are the evidence, not the absolute times).  `full`: the metrics B200_PROFILING.md names, per profiled launch,
from an `ncu --set full` report (needs the ncu binary; read-only on the report).
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("aesr::", "").replace("<unnamed>::", "")


def launches(src, dst):
    txt = open(src).read().split("\n")
    i = [k for k, l in enumerate(txt) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(txt[i:]))))
    agg = collections.OrderedDict()
    for r in rows:
        k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if not k[0].startswith("at::") and "elementwise_kernel" not in k[0])
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: `%s`, %d launches, %.1f us total, %.1f%% in aesr kernels\n\n" % (src, len(rows), tot,
                                                                                        100 * ours / tot))
        f.write("| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f%% |\n" % (k[0], k[1], k[2], a[0], a[1], a[1] / a[0],
                                                                       100 * a[1] / tot))


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEEP]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary\n\nsource: `%s`\n\n" % src)
        for r in rows[2:]:
            f.write("## launch %s: `%s`\n\n| metric | value | unit |\n|---|---|---|\n" %
                    (r[hdr.index("ID")], short(r[hdr.index("Kernel Name")])))
            for i in idx:
                f.write("| %s | %s | %s |\n" % (hdr[i], r[i], units[i]))
            f.write("\n")
    # machine-readable DRAM traffic per launch (bench.py's roofline.traffic)
    def mb(r, name):
        i = hdr.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(r[i]) * scale
    launches = [{"kernel": short(r[hdr.index("Kernel Name")]), "dram_bytes": mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum"),
                 "us": float(r[hdr.index("gpu__time_duration.sum")]),
                 "tensor_pct": float(r[hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")])}
                for r in rows[2:]]
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench            # csrc_hash(): stamps the capture to the kernel sources it was taken from
    with open(dst.replace(".md", ".json"), "w") as f:
        json.dump({"source": src, "csrc_hash": bench.csrc_hash(), "workload": os.environ.get("AESR_PROFILE_WORKLOAD", "acdc"),
                   "mean_dram_bytes_per_launch": sum(l["dram_bytes"] for l in launches) / len(launches),
                   "launches": launches}, f, indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
