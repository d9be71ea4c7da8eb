"""Per-source-line executed-instruction shares from an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n] [file_filter]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
col = rows[hdr].index("Instructions Executed")
fname, tot, out = "", 0, []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > col and r[0].strip().isdigit() and r[2] == "-":
        try:
            n = int(r[col])
        except ValueError:
            continue
        tot += n
        out.append((n, fname, int(r[0]), r[1].strip()[:120]))
out.sort(reverse=True)
print("total warp instructions", tot)
for n, f, l, s in out[:top]:
    print(f"{n / max(tot, 1) * 100:5.1f}% {n:>11d} {f}:{l}: {s}")
