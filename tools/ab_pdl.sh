#!/bin/bash
# GPU check of one build (run through gpurun): parity tests first, then the bench per launch-overlap level (MFT_PDL) and shape.
# usage: tools/ab_pdl.sh "<levels>" "<shapes>"   (defaults: "2" "5w20s 5w5s 5w50c")
LEVELS=${1:-2}
SHAPES=${2:-"5w20s 5w5s 5w50c"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for lvl in $LEVELS; do
  for shape in $SHAPES; do
    MFT_PDL=$lvl timeout 300 python bench.py --steps 200 --warmup 5 --shape $shape --no-cpu-baseline > gpurun_out/bench_pdl${lvl}_${shape}.json 2> gpurun_out/bench_pdl${lvl}_${shape}.err
    echo "pdl=$lvl $shape rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_pdl${lvl}_${shape}.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])" 2>&1 | tail -1)"
  done
done
