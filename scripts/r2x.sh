CLIPS=64 timeout 600 ncu --set full --clock-control none -k regex:conv_tsw_kernel --launch-skip 59 --launch-count 5 -f -o gpurun_out/r2_ncu_tsw3 python scripts/bench_stage.py generator conv_tsw > gpurun_out/r2_ncu_tsw3.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_ncu_tsw3.ncu-rep
