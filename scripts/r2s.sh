timeout 1200 python scripts/fp32x_e2e_probe.py > gpurun_out/r2s_fp32x_probe.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2s_fp32x_probe.log
