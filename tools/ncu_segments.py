"""Summarise an ncu source-page CSV (--page source --csv --print-source sass) by barrier-delimited SASS segments."""
import csv, re, sys
src, raw, units = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows=list(csv.reader(open(raw)))
hdr=rows[0]; unit=rows[1]; v=rows[2]
want=re.compile(r"gpu__time_duration.sum|smsp__issue_active.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread$|smsp__inst_executed.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|dram__bytes_(read|write).sum$")
for h,u,x in zip(hdr,unit,v):
    if want.search(h): print(h,x,u)
rows=list(csv.reader(open(src)))
hdr=rows[1]; data=rows[2:]
iS=hdr.index("# Samples"); iI=hdr.index("Instructions Executed"); iSrc=hdr.index("Source")
tot_s=sum(int(r[iS] or 0) for r in data); tot_i=sum(int(r[iI] or 0) for r in data)
seg=0; acc_s=0; acc_i=0; start=0; ops={}
for n,r in enumerate(data):
    s=int(r[iS] or 0); i=int(r[iI] or 0)
    acc_s+=s; acc_i+=i
    op=r[iSrc].split()[0] if not r[iSrc].startswith("@") else r[iSrc].split()[1]
    ops[op.split('.')[0]]=ops.get(op.split('.')[0],0)+i
    if "BAR.SYNC" in r[iSrc] or n==len(data)-1:
        top=sorted(ops.items(), key=lambda x:-x[1])[:5]
        if acc_i or acc_s: print(f"seg {seg} sass {start}-{n}: smp {100*acc_s/tot_s:5.1f}% inst {100*acc_i/tot_i:5.1f}% ({acc_i/units:.0f}/unit)", top)
        seg+=1; acc_s=acc_i=0; start=n+1; ops={}
