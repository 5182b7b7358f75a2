python -m pytest tests -m gpu -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2u_tests.log | tail -6
python bench.py --mode fp32 --clips 64 --steps 3 --warmup 3 --detail-out gpurun_out/r2u_detail_fp32.json > gpurun_out/r2u_bench_fp32.json 2> gpurun_out/r2u_fp32.err; echo "fp32 rc=$?"; tail -2 gpurun_out/r2u_fp32.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_detail_fp32.json'))
print(round(d['line']['value'],1), round(d['line']['e2e']['value'],1), round(d['line']['ms_per_step'],1), d['line']['clocks'].get('sm_mhz'))
for k in d['kernels'][:10]: print("  %-22s %8.2f ms %5.1f%% %8.1f %s" % (k['name'], k['ms_per_step'], 100*k['share'], k.get('achieved',0), k.get('unit','')))
for l in d['layers'][:14]: print('     ', l)
PY
