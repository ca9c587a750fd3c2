# EDT parity tests + A/B bench of an environment switch (usage: tools/quick_edt.sh VAR)
timeout 600 python -m pytest tests/test_gpu_edt.py -m gpu -x -q 2>&1 | tail -5
bash tools/ab_env.sh ${1:-MFISH_NONE}
