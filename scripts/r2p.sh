timeout 900 python -m pytest tests/test_gpu_vq.py tests/test_gpu_config_size.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2p_tests.log | tail -8
python bench.py --mode fp32 --clips 32 --steps 3 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2p_detail_fp32.json > gpurun_out/r2p_bench_fp32.json 2> gpurun_out/r2p_fp32.err; echo "fp32 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_detail_fp32.json'))
print(round(d['line']['value'],1), [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels'][:6]])
PY
