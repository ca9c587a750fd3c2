# filter parity tests + A/B bench of an environment switch (usage: tools/quick_filters.sh VAR)
timeout 600 python -m pytest tests/test_gpu_filters.py tests/test_gpu_spots.py -m gpu -x -q 2>&1 | tail -5
bash tools/ab_env.sh ${1:-MFISH_STRIDED_TMA}
