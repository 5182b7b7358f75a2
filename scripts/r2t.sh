timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fp32" > gpurun_out/r2t_ops.log 2>&1; echo "ops fp32 rc=$?"; grep -v "^$" gpurun_out/r2t_ops.log | tail -8
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -s -k "fp32" > gpurun_out/r2t_e2e.log 2>&1; echo "e2e fp32 rc=$?"; grep -v "^$" gpurun_out/r2t_e2e.log | tail -12
for tc in 1 0; do
python bench.py --mode fp32 --clips 32 --steps 3 --warmup 3 --no-cpu-baseline --opt fp32_tc=$tc --detail-out gpurun_out/r2t_detail_fp32_tc$tc.json > gpurun_out/r2t_bench_fp32_tc$tc.json 2> gpurun_out/r2t_fp32_tc$tc.err; echo "fp32 tc=$tc rc=$?"; tail -2 gpurun_out/r2t_fp32_tc$tc.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2t_detail_fp32_tc$tc.json'))
print('tc=$tc', round(d['line']['value'],1), round(d['line']['ms_per_step'],1), [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels'][:6]])
PY
done
