timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_slab.py --shape 4096 2048 200 --spots 160000 --nuclei 1000 2>/dev/null | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29532 tools/bench_slab.py --shape 4096 2048 200 --spots 160000 --nuclei 1000 2>/dev/null | tail -1
