#!/bin/bash
# ncu evidence for one benchmark step of a BASELINE config (run under gpurun; one ncu session per call).
#   tools/ncu_step.sh <tag> <profile_step.py arguments...>
# Writes gpurun_out/<tag>_metrics.csv (every kernel of the step: time, DRAM bytes, DRAM %, tensor-pipe %, L2 hit rate)
# and gpurun_out/<tag>_top_full_raw.csv (--set full of three launches of the row kernel). The .ncu-rep files stay in /tmp.
set -u
TAG=$1; shift
P="python tools/profile_step.py $*"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_utchmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,launch__registers_per_thread,launch__grid_size,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
$P > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_metrics.csv $P > /tmp/${TAG}_ncu.log 2>&1
if [ "${FULL:-1}" = "1" ]; then
  ncu --set full --clock-control none --profile-from-start off -k regex:conv_rows -c 3 -f -o /tmp/${TAG}_top $P > /tmp/${TAG}_ncu2.log 2>&1
  ncu -i /tmp/${TAG}_top.ncu-rep --page raw --csv > gpurun_out/${TAG}_top_full_raw.csv 2>/dev/null
fi
tail -2 gpurun_out/${TAG}_plain.log
wc -l gpurun_out/${TAG}_metrics.csv
