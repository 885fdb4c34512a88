tag=$1; out=gpurun_out/$tag; mkdir -p $out
export X3D_BARRIER_TIMEOUT_S=20
(time timeout 300 python -m pytest tests/test_transpose_gpu.py -x -q -k "multi_gpu and 2") > $out/pytest.log 2>&1; tail -6 $out/pytest.log
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.json; tail -3 $out/bench.err
