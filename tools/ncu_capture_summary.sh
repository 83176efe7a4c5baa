#!/bin/bash
# ncu --set full capture of one kernel, summarised to text on the GPU box (the .ncu-rep is deleted unless KEEP_REP=1):
#   tools/ncu_capture_summary.sh <out-prefix> <kernel-regex> <launch-skip> <launch-count> <command...>
out=$1; k=$2; skip=$3; cnt=$4; shift 4
ncu --set full --import-source on --clock-control none -k "regex:$k" -s "$skip" -c "$cnt" -f -o "$out" "$@" > "$out.log" 2>&1
{
  echo "# ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c $cnt $*"
  ncu -i "$out.ncu-rep" --page details 2>/dev/null | grep -E "^  [a-zA-Z_:<>]|Duration|DRAM Throughput|Memory Throughput|L2 Cache Throughput|Compute \(SM\)|Executed Ipc Active|Issue Slots Busy|Mem Busy|Max Bandwidth|L1/TEX Hit|L2 Hit|Warp Cycles Per Issued|Registers Per|Dynamic Shared|Theoretical Occ|Achieved Occ"
  ncu -i "$out.ncu-rep" --page raw --csv 2>/dev/null | python3 -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
h, u = rows[0], rows[1]
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('raw:', r[h.index('Kernel Name')][:60], '; '.join(f'{w}={r[h.index(w)]} {u[h.index(w)]}' for w in want if w in h))
"
} > "$out.txt"
[ "$KEEP_REP" = "1" ] || rm -f "$out.ncu-rep"
rm -f "$out.log"
