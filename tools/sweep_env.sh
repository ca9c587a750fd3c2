# sweep one environment variable over values on the C2 bench: tools/sweep_env.sh VAR v1 v2 ...
V=$1; shift
for val in "$@"; do
  env $V=$val python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-batch --e2e-steps 1 > gpurun_out/sw.json 2> gpurun_out/sw.err; echo "$V=$val rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/sw.json'))
print('  ms',round(d['ms_per_step'],3),' '.join(f"{k}={v['ms']:.3f}" for k,v in d['kernels'].items() if k.startswith('conv') or k.startswith('edt') or k.startswith('masked')))
PY
done
