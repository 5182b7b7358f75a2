CLIPS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tsw_kernel --launch-skip 39 --launch-count 3 -f -o gpurun_out/r2_ncu_top_final python scripts/bench_stage.py generator conv_tsw > gpurun_out/r2_ncu_top_final.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_ncu_top_final.ncu-rep
