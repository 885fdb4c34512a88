#!/bin/bash
# usage (under gpurun --gpus N): bash tools/multi_gpu_check.sh TAG N [extra bench args]
tag=$1; n=$2; shift 2
out=gpurun_out/$tag
mkdir -p $out
run() { name=$1; shift; (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 "$@") > $out/$name.json 2> $out/$name.err; tail -c 700 $out/$name.json; tail -3 $out/$name.err; }
if [ "$n" = "2" ]; then
  (time timeout 600 python -m pytest tests/test_transpose_gpu.py -x -q -k "multi_gpu and 2") > $out/pytest.log 2>&1; tail -4 $out/pytest.log
fi
run bench "$@"
X3D_OVERLAP=2 run bench_overlap2 --no-e2e --no-cpu-baseline "$@"
