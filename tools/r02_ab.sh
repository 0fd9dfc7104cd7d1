#!/bin/bash
# A/B of an environment switch: tools/r02_ab.sh VAR "v1 v2 ..." [shape]
VAR=$1; VALS=$2; SHAPE=${3:-5w20s}
mkdir -p gpurun_out
for v in $VALS; do
  env $VAR=$v timeout 300 python bench.py --steps 200 --warmup 5 --shape $SHAPE --no-cpu-baseline --no-gpu-reference > gpurun_out/ab_${VAR}_${v}_$SHAPE.json 2> gpurun_out/ab_${VAR}_${v}_$SHAPE.err || tail -5 gpurun_out/ab_${VAR}_${v}_$SHAPE.err
  python -c "import json;d=json.load(open('gpurun_out/ab_${VAR}_${v}_$SHAPE.json'));print('$VAR=$v $SHAPE ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['kernel_ms_per_step'])"
done
