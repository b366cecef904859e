#!/bin/bash
# usage: scripts/ncu_summary.sh report.ncu-rep kernel-regex mangled-substring [top]
# prints key raw metrics of the first matching launch, then the per-source-line aggregation
REP=$1; KRE=$2; SUB=$3; TOP=${4:-30}
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); hdr=r[0]; units=r[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors_srcunit_tex_op_read.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(f'{w:90s} {units[i]:10s}', [row[i] for row in r[2:]])
"
ncu -i $REP --page source --csv --kernel-name regex:$KRE --launch-count 1 2>/dev/null > /tmp/ncu_sass.csv
mkdir -p /tmp/dis && (cd /tmp/dis && rm -f *.cubin && cuobjdump -xelf all /root/repo/mcmc_spec_b200/libmcmcspec.so >/dev/null && nvdisasm -g -c *.cubin > dis.txt 2>&1)
python /root/repo/scripts/ncu_by_line.py /tmp/ncu_sass.csv /tmp/dis/dis.txt "$SUB" $TOP
