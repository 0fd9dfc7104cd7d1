#!/bin/bash
# One `ncu --set full --import-source on` capture each of the 192->192 forward and dgrad launches of
# the 5w20s head (second Wcompute: full 89 040 rows), eager launches (no graph), for --page source reading.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'umma_rows_kernel<mft::BnActT' -s 3 -c 1 \
    -o gpurun_out/src_fwd -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_fwd.log 2>&1
echo "fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'DhInPlaceT, mft::EpiDyU' -s 2 -c 1 \
    -o gpurun_out/src_dgrad -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_dgrad.log 2>&1
echo "dgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
