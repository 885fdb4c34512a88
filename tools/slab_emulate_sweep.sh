for nv in 2 4 8; do
  X3D_SLABZ_EMULATE=$nv timeout 200 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2j/emu$nv.json 2> gpurun_out/r2j/emu$nv.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2j/emu$nv.json').read().strip().splitlines()[-1])
print($nv, d['ms_per_step'], [(c['name'][:32],c['count'],round(c['avg_ms'],3)) for c in d['roofline']['classes'] if 'momentum' in c['name']])
P
done
