import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[idx[k]].replace(',',''))
    except Exception: return float('nan')
stalls=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
agg={}
for r in rows[2:]:
    name=r[idx['Kernel Name']].split('(')[0][:26]
    a=agg.setdefault(name,{'n':0,'t':0.0,'rows':[]}); a['n']+=1; a['t']+=f(r,'gpu__time_duration.sum'); a['rows'].append(r)
print('time unit', rows[1][idx['gpu__time_duration.sum']])
for name,a in sorted(agg.items(), key=lambda kv:-kv[1]['t']):
    r=max(a['rows'], key=lambda r: f(r,'gpu__time_duration.sum'))
    st=sorted(((f(r,s), s.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for s in stalls), reverse=True)[:4]
    print(f"{name:26s} n={a['n']:2d} t={a['t']:9.1f} issue%={f(r,'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):5.1f} inst%={f(r,'sm__inst_executed.avg.pct_of_peak_sustained_elapsed'):5.1f} warps%={f(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} dram%={f(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} lat/inst={f(r,'smsp__average_warp_latency_per_inst_issued.ratio'):5.1f} top:{[(round(v,1),n) for v,n in st]}")
