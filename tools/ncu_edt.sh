# ncu --set full of the two envelope kernels (x, z) of the anisotropic C2 transform -> gpurun_out/<tag>.ncu-rep
TAG=${1:-r2_edt}
timeout 300 python tools/bench_edt.py --child prof > gpurun_out/plain_edt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:edt_hull\|edt_envelope -s 6 -c 2 -f -o gpurun_out/$TAG python tools/bench_edt.py --child prof > gpurun_out/ncu_edt.log 2>&1; echo ncu=$?; tail -3 gpurun_out/ncu_edt.log
