# spot + EDT parity files of the GPU suite, then the bench with the per-kernel table
timeout 900 python -m pytest tests/test_gpu_spots.py tests/test_gpu_edt.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_b.json 2> gpurun_out/bench_c2_b.err; echo bench=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c2_b.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k,v in d['kernels'].items(): print(f"  {k:24s} {v['ms']:.3f} ms  frac {v.get('frac',0):.3f}")
PY
