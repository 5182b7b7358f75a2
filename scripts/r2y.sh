python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_modules.py tests/test_gpu_config_size.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt post_tc=$v --detail-out gpurun_out/r2y_detail_pt${v}.json > gpurun_out/r2y_pt${v}.json 2> gpurun_out/r2y_pt${v}.err
    python - <<PY
import json
l=json.loads(open('gpurun_out/r2y_pt${v}.json').read().strip().splitlines()[-1])
d=json.load(open('gpurun_out/r2y_detail_pt${v}.json'))
k={x['name']:(x['ms_per_step'],x.get('frac')) for x in d['kernels']}
print('post_tc=$v', round(l['ms_per_step'],1), l['clocks']['sm_mhz'], l.get('codes_checksum'), l['gpu_launches'], k.get('conv_post_tanh'))
PY
done
