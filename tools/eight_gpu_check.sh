#!/bin/bash
# usage (under gpurun --gpus 8): bash tools/eight_gpu_check.sh TAG [worker]
# the 512^3 bench line on 8 GPUs and BASELINE config #5 (1536^3 on 8 GPUs); with "worker" also the multi-GPU worker test on 8 ranks
tag=$1
out=gpurun_out/$tag
mkdir -p $out
export X3D_BARRIER_TIMEOUT_S=20   # a rank that never arrives is reported after 20 s instead of 120
run() { n=$1; name=$2; shift 2; (time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@") > $out/$name.json 2> $out/$name.err; tail -c 900 $out/$name.json; tail -3 $out/$name.err; }
if [ "$2" = "worker" ]; then
  (time timeout 250 python -m pytest tests/test_transpose_gpu.py -x -q -k "multi_gpu and 8") > $out/pytest.log 2>&1; tail -15 $out/pytest.log
  grep -q "1 passed" $out/pytest.log || { echo "worker test failed: not running the benches"; exit 1; }
fi
run 8 bench8 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline
run 8 bench8_1536 --size 1536 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline
