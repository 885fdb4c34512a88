echo "== tests"; timeout 260 python -m pytest tests/test_transpose_gpu.py -x -q -k multi 2>&1 | tail -4
echo "== bench overlap"; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1y_2gpu.json 2>gpurun_out/bench_r1y_2gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r1y_2gpu.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["config"].get("diagnostics_after_run"))
    for c in d["roofline"]["classes"]:
        if "transpose" in c["name"] or "momentum" in c["name"] or "elementwise" in c["name"]: print(c)
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r1y_2gpu.err").read()[-1500:])
PY
