#!/bin/bash
# usage (under gpurun --gpus N): bash tools/multi_gpu_check.sh TAG N [extra bench args]
# the multi-GPU worker test (N = 2 only), the bench line of the default configuration, and the same with the slab z kernels
# off (X3D_SLABZ=0: the z part of the momentum terms through y <-> z transposes) for comparison
tag=$1; n=$2; shift 2
out=gpurun_out/$tag
mkdir -p $out
run() { name=$1; shift; (time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 "$@") > $out/$name.json 2> $out/$name.err; tail -c 1200 $out/$name.json; tail -3 $out/$name.err; }
if [ "$n" = "2" ]; then
  (time timeout 400 python -m pytest tests/test_transpose_gpu.py -x -q -k "multi_gpu and 2") > $out/pytest.log 2>&1; tail -15 $out/pytest.log
fi
run bench "$@"
X3D_SLABZ=0 run bench_transposes --no-e2e --no-cpu-baseline "$@"
