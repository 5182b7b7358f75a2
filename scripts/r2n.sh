timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2n_ops.log 2>&1; echo "ops rc=$?"; grep -v "^$" gpurun_out/r2n_ops.log | tail -12
timeout 1200 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/r2n_e2e.log 2>&1; echo "e2e rc=$?"; tail -4 gpurun_out/r2n_e2e.log
for cl in 2 1 2 1; do
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt cta_pairs=$cl --detail-out gpurun_out/r2n_detail_cl$cl.json > gpurun_out/r2n_bench_cl$cl.json 2> gpurun_out/r2n_bench_cl$cl.err; echo "bench cl=$cl rc=$?"; tail -2 gpurun_out/r2n_bench_cl$cl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2n_detail_cl$cl.json'))
print('cl=$cl', round(d['line']['ms_per_step'],1), d['line']['clocks'].get('sm_mhz'), [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels'] if ('conv_ts' in k['name'])])
PY
done
