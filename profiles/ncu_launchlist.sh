#!/bin/bash
# One step of the bench workload under ncu (launch list with the per-kernel counters launchsum.py prints):
#   profiles/ncu_launchlist.sh <out.csv> [skip] [count]      (run only after the same bench command exited 0 without ncu)
out=${1:-gpurun_out/launches.csv}; skip=${2:-28}; count=${3:-7}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
M=$M,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -s $skip -c $count --csv --log-file $out \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launchlist.log 2>&1
