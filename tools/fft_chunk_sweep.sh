mkdir -p gpurun_out/r2g
for ch in 0 2 4 16 32 64; do
  X3D_FFT_CHUNK=$ch python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2g/bench_chunk$ch.json 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2g/bench_chunk$ch.json").read().strip().splitlines()[-1])
print("chunk $ch", round(d["ms_per_step"],3), [(c["name"][:24], round(c["avg_ms"],3)) for c in d["roofline"]["classes"] if c["name"].startswith(("fft","poisson"))])
PY
done
echo "== solver tests with the dual solves"
python -m pytest tests/test_solver_gpu.py tests/test_cases_gpu.py -x -q 2>&1 | tail -3
