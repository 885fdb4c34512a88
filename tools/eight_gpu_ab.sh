#!/bin/bash
# usage (under gpurun --gpus 8): bash tools/eight_gpu_ab.sh TAG -- 512^3 on 8 GPUs with and without X3D_OVERLAP_DIV
tag=$1; out=gpurun_out/$tag; mkdir -p $out
export X3D_BARRIER_TIMEOUT_S=20
for v in 1 0; do
  (time X3D_OVERLAP_DIV=$v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline) > $out/bench8_ovl$v.json 2> $out/bench8_ovl$v.err
  python - <<P
import json
d=json.loads(open('$out/bench8_ovl$v.json').read().strip().splitlines()[-1])
print('overlap_div=$v', d['ms_per_step'], d['clocks'], [(c['name'][:28],c['count'],round(c['total_ms'],3)) for c in d['roofline']['classes'] if 'transpose' in c['name'] or 'ring' in c['name']])
P
done
