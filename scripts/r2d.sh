python -m pytest tests/test_gpu_config_size.py tests/test_gpu_dropin_joined.py tests/test_gpu_modules.py -m gpu -q -s > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2d_tests.log | tail -25
for m in 0 1 2 3 4 8 16 32 64 7 15; do
  echo "== DC_PAIRX_DBG=$m"; DC_PAIRX_DBG=$m CLIPS=128 python scripts/bench_stage.py generator pairx 2>&1 | tail -10
done > gpurun_out/r2d_probe.log 2>&1
grep "==\|total" gpurun_out/r2d_probe.log
