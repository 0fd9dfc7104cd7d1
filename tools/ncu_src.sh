#!/bin/bash
# One `ncu --set full --import-source on` capture of a tcgen05 kernel of the 5w20s head (eager launches, no graph),
# for `ncu --page source --csv` + tools/ncu_regions.py.   usage: tools/ncu_src.sh <demangled-name regex> <skip> <out>
PAT=${1:-"umma_wgrad_kernel<mft::DhT, mft::BnActQT, 6, 6>"}
SKIP=${2:-1}
OUT=${3:-src_wgrad}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$PAT" -s $SKIP -c 1 \
    -o gpurun_out/$OUT -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_$OUT.log 2>&1
echo "rc=$?"; ls -la gpurun_out/$OUT.ncu-rep
