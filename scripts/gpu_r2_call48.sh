#!/bin/bash
mkdir -p gpurun_out
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.tc.so
timeout 300 ncu --set full --clock-control none -k regex:tree_attn_tc -c 2 -o /tmp/attn_tc -f python scripts/ncu_attn_tc.py > /dev/null 2>&1
ncu -i /tmp/attn_tc.ncu-rep --page raw --csv 2>/dev/null > /tmp/attn_tc.csv
python - <<'P' > gpurun_out/r2_ncu_attn_tc.txt
import csv
rows=list(csv.reader(open('/tmp/attn_tc.csv')))
hdr=rows[0]
d=dict(zip(hdr,rows[-1]))
keys=['gpu__time_duration.sum','dram__bytes_read.sum','lts__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.sum','smsp__warps_eligible.avg.per_cycle_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_xu.sum']
for k in keys: print(k, d.get(k))
for h,v in d.items():
    if 'issue_stalled' in h and 'ratio' in h:
        try:
            if float(v)>0.2: print(' ', h.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_',''), v)
        except: pass
P
cat gpurun_out/r2_ncu_attn_tc.txt
