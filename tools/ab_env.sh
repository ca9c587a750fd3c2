# A/B one environment switch on the C2 bench: tools/ab_env.sh VAR
V=$1
for val in 0 1; do
  env $V=$val python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$val.json 2> gpurun_out/ab_$val.err; echo "$V=$val rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$val.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k,v in d['kernels'].items(): print(f"  {k:24s} {v['ms']:.3f} ms")
PY
done
