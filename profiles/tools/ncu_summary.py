"""Per-kernel summary of an ncu --set full report: time, issue/dram utilisation, occupancy, top stalls, DRAM bytes."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return float('nan')
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
print("| # | kernel | us | issue% | dram% | DRAM MB (rd+wr) | occ% (achieved/theor) | regs | warp lat/inst | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n, r in enumerate(rows[2:]):
    name = r[ix['Kernel Name']].split('(')[0]
    st = sorted(((f(r, s), s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for s in stalls if f(r, s) == f(r, s)), reverse=True)[:4]
    mb = (f(r, 'dram__bytes_read.sum') + f(r, 'dram__bytes_write.sum'))
    unit = r[ix['dram__bytes_read.sum']]
    u = rows[1][ix['dram__bytes_read.sum']]
    mb *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)
    tu = rows[1][ix['gpu__time_duration.sum']]
    t = f(r, 'gpu__time_duration.sum') * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(tu, 1.0)
    print(f"| {n} | {name} | {t:.1f} | {f(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{f(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {mb:.1f} | "
          f"{f(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}/{f(r,'sm__maximum_warps_per_active_cycle_pct'):.0f} | "
          f"{f(r,'launch__registers_per_thread'):.0f} | {f(r,'smsp__average_warp_latency_per_inst_issued.ratio'):.1f} | "
          f"{', '.join(f'{n_} {v:.1f}' for v, n_ in st)} |")
