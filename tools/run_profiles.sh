# round-stamped evidence: bench line, ncu launch list of the same command, ncu full capture of the main kernels
R=${1:-r2}
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_c2.json 2> gpurun_out/${R}_bench_c2.err; echo bench=$?
LIST="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batch --quick --e2e-steps 1"
$LIST > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv_|nms_|edt_|qsel_|minmax|ingest_|blur_|gather_|prune_|iota|desc_key|apply_rank|RadixSort|morph_' -c 600 --csv --log-file gpurun_out/${R}_launches_c2.csv $LIST > gpurun_out/ncu.log 2>&1; echo ncu_list=$?
python tools/prof_step.py C2 > gpurun_out/plain_prof.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:edt_envelope|nms_verify|conv_strided|conv_y_kernel|edt_scan|qsel_range|qsel_hist' -s 14 -c 14 -f -o gpurun_out/${R}_prof_c2 python tools/prof_step.py C2 > gpurun_out/ncu_full.log 2>&1; echo ncu_full=$?
