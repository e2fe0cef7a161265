"""Per-CUDA-source-line instruction counts and stall samples for one kernel launch of an ncu report.
usage: ncu_lines.py report.ncu-rep <launch-index> [top]"""
import csv, subprocess, sys, io
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass --launch-skip {kid} --launch-count 1",
                     shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
print(rows[1][1][:100])
lines = []
for r in rows[hi + 1:]:
    if r and r[0].isdigit():
        try: lines.append((int(r[0]), r[1], int(r[i_i]), int(r[i_s])))
        except ValueError: pass
ti = sum(l[2] for l in lines); ts = sum(l[3] for l in lines)
print("total warp-inst", ti, "samples", ts)
if "--order" in sys.argv:
    for ln, src, i, s in lines:
        if i * 200 > ti or s * 200 > ts:
            print(f"{ln:4d} {100*i/ti:5.1f}% inst {100*s/ts:5.1f}% smp | {src.strip()[:120]}")
else:
    for ln, src, i, s in sorted(lines, key=lambda l: -l[2])[:top]:
        print(f"{ln:4d} {100*i/ti:5.1f}% inst {100*s/ts:5.1f}% smp | {src.strip()[:120]}")
