python profiles/ncu_case.py cleanup5 65536 > /dev/null
ncu --set full --clock-control none --import-source on -k regex:ssd_kernel -s 30 -c 1 -f -o gpurun_out/stall python profiles/ncu_case.py cleanup5 65536 > /dev/null 2>&1
ncu -i gpurun_out/stall.ncu-rep --page raw --csv > gpurun_out/stall_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/stall_raw.csv")))
h,u,v=rows[0],rows[1],rows[2]
out=[]
for i,name in enumerate(h):
    if any(k in name for k in ("issue_stalled","pipe_","inst_executed_pipe","warps_eligible","issue_active","l1tex__data_bank","lsu_mem_shared","smsp__inst_executed_op","warp_issue_stalled")) and "pct" in name or "per_warp_active" in name or "ratio" in name and "stalled" in name:
        out.append((name,u[i],v[i]))
for n,uu,vv in sorted(out): print("%-90s %12s %s"%(n,vv,uu))
PY
rm -f gpurun_out/stall.ncu-rep gpurun_out/stall_raw.csv
