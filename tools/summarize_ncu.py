"""Turns ncu outputs brought back from the GPU box into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_kernels.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0][:90]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path, errors="ignore") if l.startswith('"'))]
    hdr = rows[0]
    iname, ival, imet = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    iunit = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[imet] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        v = v / 1e3 if r[iunit] in ("ns", "nsecond") else (v * 1e3 if r[iunit] in ("ms", "msecond") else v)   # -> us
        a = agg[short(r[iname])]
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / total:.1f}% |\n")
        f.write(f"\ntotal {total / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches (ncu-serialised, cold caches: compare shares, not absolutes)\n")


METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
           "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    iname = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    with open(out, "w") as f:
        f.write("| kernel | " + " | ".join(m for m, _ in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
        f.write("| (unit) | " + " | ".join(units[i] for _, i in cols) + " |\n")
        for r in rows[2:]:
            f.write(f"| `{short(r[iname])}` | " + " | ".join(r[i] for _, i in cols) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
