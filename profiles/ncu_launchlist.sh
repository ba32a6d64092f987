#!/bin/bash
# One step of the bench workload under ncu (launch list with the per-kernel counters launchsum.py prints):
#   profiles/ncu_launchlist.sh <out.csv> [skip] [count]   (skip: launches before the step that is listed -- the one-off
#   shared-memory probe of tq_create + 4 steps of 7 launches = 29; the committed r02 list was taken with 28 and therefore starts
#   with the reduce launch of the step before: the same seven kernels of a steady-state step)      (run only after the same bench command exited 0 without ncu)
out=${1:-gpurun_out/launches.csv}; skip=${2:-29}; count=${3:-7}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
M=$M,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -s $skip -c $count --csv --log-file $out \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launchlist.log 2>&1
