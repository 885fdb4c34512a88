#!/bin/bash
# One GPU box call at the end of a round: GPU test suite, the bench line, the ncu launch list of the same command and
# a counter capture of the fused momentum kernels.  Usage: gpurun --timeout 720 -- 'bash tools/gpu_round_check.sh r1z'
tag=${1:-run}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 330 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
echo "== bench"
timeout 200 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -c 400 gpurun_out/bench_${tag}.json
echo "== ncu launch list"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_${tag}.log 2>&1
wc -l gpurun_out/launches_${tag}.csv
echo "== ncu fused momentum kernels"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
timeout 100 ncu --metrics $M --clock-control none -k regex:k_mom_pair -c 3 --csv --log-file gpurun_out/mom_${tag}.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_mom_${tag}.log 2>&1
wc -l gpurun_out/mom_${tag}.csv
