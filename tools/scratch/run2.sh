for m in 2 1; do
  echo "== mode $m tests"; X3D_P2P_MODE=$m timeout 200 python -m pytest tests/test_transpose_gpu.py -x -q -k multi 2>&1 | tail -4
  echo "== mode $m bench"; X3D_P2P_MODE=$m timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$m bench.py --gpus 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1x_2gpu_mode$m.json 2>gpurun_out/bench_r1x_2gpu_mode$m.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r1x_2gpu_mode$m.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"])
    for c in d["roofline"]["classes"]:
        if "transpose" in c["name"] or "momentum" in c["name"]: print(c)
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r1x_2gpu_mode$m.err").read()[-1500:])
PY
done
