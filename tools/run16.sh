python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e']['value'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/bench_slab.py --steps 3 --warmup 1 2> gpurun_out/slab8.err | tail -1 > gpurun_out/slab_n8_c4.json; cat gpurun_out/slab_n8_c4.json; tail -2 gpurun_out/slab8.err
