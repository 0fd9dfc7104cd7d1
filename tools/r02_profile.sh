#!/bin/bash
# ncu evidence of the round: launch list of one eager step + DRAM traffic / tensor activity of the GEMM kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__cycles_active.avg \
    --clock-control none -k regex:"umma_" -s 123 -c 82 --csv --log-file gpurun_out/r02_gemm_metrics.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_ncu_metrics.log 2>&1
echo "metrics rc=$?"
