N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 tools/bench_slab.py --steps 3 --warmup 1 2> gpurun_out/slab$N.err | tail -2 > gpurun_out/slab_n${N}_c4.json; cat gpurun_out/slab_n${N}_c4.json; tail -2 gpurun_out/slab$N.err | cut -c1-300
