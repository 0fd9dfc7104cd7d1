#!/bin/bash
# End-of-round measurement on ONE B200: parity tests, the bench matrix, the reference arm, the ncu launch
# list of the default command and one ncu capture (time, DRAM bytes, tensor/issue activity) of every
# tcgen05 / row kernel of one eager step.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
MFT_PDL=0 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_pdl0.log 2>&1; echo "pytest(MFT_PDL=0) rc=$?"; tail -1 $O/pytest_gpu_pdl0.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
run() { name=$1; shift; timeout 600 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json; d=json.load(open('$O/$name.json')); print(d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)"; }
run r01_final_bench
run r01_final_noshare --no-share --no-cpu-baseline
run r01_final_5w5s --shape 5w5s --no-cpu-baseline
run r01_final_5w50c --shape 5w50c --no-cpu-baseline
run r01_final_fp32 --precision fp32 --steps 100 --no-cpu-baseline
MFT_PDL=0 run r01_final_pdl0 --no-cpu-baseline
MFT_PDL=0 run r01_final_pdl0_5w5s --shape 5w5s --no-cpu-baseline
run r01_reference_arm --impl reference --steps 3 --warmup 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file $O/r01_tf32_launches_v10.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --kernel-name-base demangled -k regex:'mft::' -s 0 -c 118 -o $O/r01_sections -f \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $O/ncu_sections.log 2>&1; echo "sections rc=$?"
ls -la $O/*.ncu-rep
