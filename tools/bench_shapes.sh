#!/bin/bash
# Kernel micro-benchmarks on the shapes SURVEY 8(d) lists: config #2 (512^3), #3 channel, #4 cylinder and the per-GPU
# shards of #5 (1536^3 on 8 GPUs).  Usage: gpurun --timeout 400 -- 'bash tools/bench_shapes.sh TAG'
tag=${1:-run}
mkdir -p gpurun_out
run() { timeout 90 python tools/bench_kernels.py --n $1 $2 $3 --ops $4 --iters 10 2>&1 | grep '^{' ; }
{
echo "# config 2: 512^3 periodic";                run 512 512 512 derx_00,dery_00,derz_00,derxx_00,deryy_00,derzz_00,filx_00,interxvp,deryvp,derzpv
echo "# config 3: channel 256x129x128 (00,22,00)"; run 256 129 128 derx_00,dery_22,derz_00,derxx_00,deryy_22,derzz_00,interxvp,deryvp,derzpv
echo "# config 4: cylinder 769x256x32 (22,00,00)"; run 769 256 32 derx_22,dery_00,derz_00,derxx_22,deryy_00,derzz_00,interxvp,deryvp,derzpv
echo "# config 5 slab shard 1536x1536x192";        run 1536 1536 192 derx_00,dery_00,derxx_00,deryy_00,interxvp,deryvp
echo "# config 5 pencil shard 1536x768x384 (x lines)"; run 1536 768 384 derx_00,derxx_00
} > gpurun_out/kernels_${tag}.txt
cat gpurun_out/kernels_${tag}.txt
