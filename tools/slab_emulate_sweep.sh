#!/bin/bash
# usage: bash tools/slab_emulate_sweep.sh TAG -- the 512^3 step on ONE GPU with its planes treated as 2 / 4 / 8 virtual slabs
# (X3D_SLABZ_EMULATE): one launch of k_mom_slab / k_zfix then costs what it costs each of 2 / 4 / 8 ranks
tag=${1:-emu}
mkdir -p gpurun_out/$tag
for nv in 2 4 8; do
  X3D_SLABZ_EMULATE=$nv timeout 200 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/$tag/emu$nv.json 2> gpurun_out/$tag/emu$nv.err
  python - <<P
import json
d=json.loads(open('gpurun_out/$tag/emu$nv.json').read().strip().splitlines()[-1])
print($nv, d['ms_per_step'], [(c['name'][:32],c['count'],round(c['avg_ms'],3)) for c in d['roofline']['classes'] if 'momentum' in c['name']])
P
done
