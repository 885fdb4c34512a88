"""bench.py contract, CPU side: the reference arm (`--impl reference`) runs the oracle port on the host cores and
prints ONE JSON line with the keys the driver reads; under torchrun only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "24", "--steps", "2",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line():
    out = _run()
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "grid-point-steps/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "24^3" in d["config"]["workload"]   # the reference arm runs the SAME workload as the GPU arm
    assert d["steps"] == 2


def test_reference_arm_ignores_torchrun_omp_num_threads():
    # torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must still use every core it is allowed
    d = json.loads(_run({"OMP_NUM_THREADS": "1"}).strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_without_work():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""
