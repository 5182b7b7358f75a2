timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_modules.py tests/test_gpu_dropin_joined.py -m gpu -x -q -s -k "fp32 or encode or decode or generator or tiny or dropin or joined" > gpurun_out/r2v_e2e.log 2>&1; echo "e2e rc=$?"; grep -v "^$" gpurun_out/r2v_e2e.log | tail -8
python bench.py --mode fp32 --clips 64 --steps 3 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2v_detail_fp32.json > gpurun_out/r2v_bench_fp32.json 2> gpurun_out/r2v_fp32.err; echo "fp32 rc=$?"; tail -2 gpurun_out/r2v_fp32.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2v_detail_fp32.json'))
print(round(d['line']['value'],1), round(d['line']['ms_per_step'],1), [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels'][:6]])
PY
