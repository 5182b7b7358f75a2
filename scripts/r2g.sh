python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2g_tests.log | tail -12
python bench.py --steps 10 --warmup 3 --detail-out gpurun_out/r2g_detail_recon_n1.json > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2g_bench_n1.json; echo
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_detail_recon_n1.json'))
for k in d['kernels']:
    if k['name'].startswith('vq'): print(k['name'], round(k['ms_per_step'],3))
PY
