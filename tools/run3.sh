python tools/diag_step.py > gpurun_out/diag.log 2>&1; tail -20 gpurun_out/diag.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
