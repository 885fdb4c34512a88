#!/bin/bash
# One GPU box call: GPU test suite, the bench line, the ncu launch list of the same command, a counter capture and a
# `--set full` capture of the dominant kernel (k_mom_pair) and of k_pair on y and z lines.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh r2x'
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest -m gpu"
(time timeout 400 python -m pytest tests -q -m gpu -x) > $out/pytest.log 2>&1
tail -4 $out/pytest.log
echo "== bench"
timeout 300 python bench.py > $out/bench.json 2> $out/bench.err
tail -c 300 $out/bench.json
echo "== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $out/ncu_bench.log 2>&1
wc -l $out/launches.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
echo "== ncu counters: k_mom_pair (z, y, x+intt of one sub-step)"
timeout 150 ncu --metrics $M --clock-control none -k regex:k_mom_pair -s 3 -c 3 --csv --log-file $out/mom.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $out/ncu_mom.log 2>&1
wc -l $out/mom.csv
echo "== ncu counters: k_pair / k_contig / k_spec (one sub-step)"
timeout 150 ncu --metrics $M --clock-control none -k regex:'k_pair|k_contig|k_spec' -s 16 -c 17 --csv --log-file $out/ops.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $out/ncu_ops.log 2>&1
wc -l $out/ops.csv
echo "== ncu --set full: k_mom_pair"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_mom_pair -s 3 -c 3 -f -o $out/mom_full \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $out/ncu_mom_full.log 2>&1
ls -la $out/*.ncu-rep
