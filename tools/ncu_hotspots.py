"""Development aid: stall reasons and per-SASS-instruction execution / sample counts of the first kernel of an ncu report
(`ncu --set full --import-source on`), grouped by execution count (= code region: once per warp, per tile, per chunk)."""
import csv,sys,subprocess,collections
path=sys.argv[1]
raw=subprocess.run(["ncu","-i",path,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); h=rows[0]; r=rows[2]
def get(n):
    return r[h.index(n)] if n in h else None
for n in ["gpu__time_duration.sum","dram__bytes_read.sum.per_second","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread"]:
    print(n, get(n))
st=[(n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), float(r[i])) for i,n in enumerate(h) if 'smsp__average_warps_issue_stalled' in n and 'ratio' in n]
print(' '.join(f"{n}={v:.2f}" for n,v in sorted(st,key=lambda x:-x[1])[:9]))
src=subprocess.run(["ncu","-i",path,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
body=[]
for rr in rows[2:]:
    if len(rr)<10 or rr[0]=='Kernel Name': break
    if rr[0]=='Address': continue
    body.append((rr[ia].strip(), int(rr[ie]), int(rr[isamp])))
tot=sum(b[2] for b in body)
g=collections.defaultdict(lambda:[0,0,0])
for s,e,sm in body:
    g[e][0]+=1; g[e][1]+=sm; g[e][2]+=e
for e,v in sorted(g.items(), key=lambda kv:-kv[1][1])[:8]: print('exec',e,'n_instr',v[0],'samples',v[1],'pct %.1f'%(100*v[1]/tot),'inst_total',v[2])
for i,(s,e,sm) in sorted(enumerate(body),key=lambda x:-x[1][2])[:12]: print(i,e,sm,s[:80])
