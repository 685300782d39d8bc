import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; data=rows[2:]
pats=sys.argv[2:] or ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit','smsp__issue_active.avg.pct','smsp__inst_executed.sum','bank_conflicts','issue_stalled.*per_issue_active','pipe_fp64','op_local','op_shared','wavefronts_mem_shared.sum','lts__t_bytes.sum','l1tex__t_bytes.sum']
import re
for d in data:
    print('-----')
    for i,h in enumerate(hdr):
        if any(re.search(p,h) for p in pats):
            v=d[i]
            if h!='Kernel Name':
                try:
                    if float(v.replace(',',''))==0: continue
                except: pass
            print(f"{h:100s} {v[:60]:>22s} {units[i]}")
