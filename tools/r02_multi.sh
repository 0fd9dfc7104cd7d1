#!/bin/bash
# multi-GPU evidence (run with gpurun --gpus N): NCCL test, weak-scaling bench with and without the 21.2 MB payload, eval arm
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q 2>&1 | tail -3
for extra in "" "--dp-bytes 21.2"; do
  tag=$(echo "$extra" | tr -d ' -.')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 100 --warmup 5 $extra > gpurun_out/r02_n${N}_train${tag}.json 2> gpurun_out/r02_n${N}_train${tag}.err
  python -c "import json;d=json.load(open('gpurun_out/r02_n${N}_train${tag}.json'));print('train', '$extra', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['parallelism'], d['clocks'])"
done
for shape in 5w5s 5w20s 5w50c; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus $N --mode eval --shape $shape > gpurun_out/r02_n${N}_eval_$shape.json 2> gpurun_out/r02_n${N}_eval_$shape.err
  python -c "import json;d=json.load(open('gpurun_out/r02_n${N}_eval_$shape.json'));print('eval', '$shape', d['n_gpus'], d['value'], d['ms_per_step'])"
done
