CLIPS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ts_kernel --launch-skip 57 --launch-count 1 -f -o gpurun_out/r2_ncu_ts1 python scripts/bench_stage.py generator conv_ts > gpurun_out/r2_ncu_ts1.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_ncu_ts1.ncu-rep
