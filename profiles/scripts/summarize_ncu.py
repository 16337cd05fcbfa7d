import csv, collections, subprocess, sys
rep=sys.argv[1]; natoms=int(sys.argv[2]); ntr=1000
det=subprocess.run(f"ncu -i {rep} --page details",shell=True,capture_output=True,text=True).stdout
for l in det.split('\n'):
    if any(k in l for k in ["Duration","Achieved Occupancy","Registers Per","Executed Ipc Active","Issue Slots Busy","L1/TEX Hit","Warp Cycles Per Issued","Theoretical Occ"]): print(l.strip())
raw=subprocess.run(f"ncu -i {rep} --page raw --csv",shell=True,capture_output=True,text=True).stdout
rows=list(csv.reader(raw.split('\n')))
hdr=rows[0]; vals=rows[2]
out=[]
for h,v in zip(hdr,vals):
    if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
        out.append((float(v), h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
    if h in ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'): print(h,v)
print(sorted(out,reverse=True)[:7])
src=subprocess.run(f"ncu -i {rep} --page source --csv",shell=True,capture_output=True,text=True).stdout
rows=list(csv.reader(src.split('\n')))
hdr=rows[1]; data=rows[2:]
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iW=hdr.index("Warp Stall Sampling (All Samples)")
tot=0; byop=collections.Counter(); samp=collections.Counter(); allsamp=0
for r in data:
    try: n=int(r[iE]); w=int(r[iW])
    except: continue
    op=r[iS].strip().split()[0]
    if op.startswith('@'): op=r[iS].strip().split()[1]
    op=op.split('.')[0]
    byop[op]+=n; samp[op]+=w; tot+=n; allsamp+=w
print("instr per TR", tot/natoms/ntr, "samples", allsamp)
for op,n in byop.most_common(18): print(f"{op:10s} {n/natoms/ntr:8.1f} per TR   samples {100*samp[op]/allsamp:5.1f}%")
