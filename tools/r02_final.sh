#!/bin/bash
# End-of-round measurement on ONE B200 (round 2): parity tests, smoke, the bench matrix, the reference arm, the
# ncu launch list of the default command, one ncu metrics capture (time, DRAM bytes, tensor / issue activity) of
# every library kernel of one eager step, and one `--set full --import-source on` capture of the top GEMM kernel.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_smoke.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json; d=json.load(open('$O/$name.json')); print(d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('frac_on_step'))" 2>&1 | tail -1)"; }
run r02_final_bench
run r02_final_driver_cmd --steps 20 --warmup 5
run r02_reference_arm --impl reference --steps 20 --warmup 5
run r02_final_noshare --no-share --no-cpu-baseline --no-gpu-reference
run r02_final_5w5s --shape 5w5s --no-cpu-baseline
run r02_final_5w50c --shape 5w50c --no-cpu-baseline
run r02_final_fp32 --precision fp32 --steps 100 --no-cpu-baseline --no-gpu-reference
MFT_BWD_SPLIT=0 run r02_final_nosplit --no-cpu-baseline --no-gpu-reference
for shape in 5w5s 5w20s 5w50c; do run r02_eval_$shape --mode eval --shape $shape; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > $O/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum \
    --clock-control none --kernel-name-base demangled -k regex:'mft::' -s 0 -c 130 -o $O/r02_sections -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > $O/r02_ncu_sections.log 2>&1; echo "sections rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'umma_rows_kernel<mft::DhInPlaceT, mft::EpiDyU>' -s 3 -c 1 -o $O/r02_src_dgrad -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > $O/r02_ncu_src.log 2>&1; echo "src rc=$?"
ls -la $O/*.ncu-rep
