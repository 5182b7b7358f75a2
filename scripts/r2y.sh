timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2y_detail_new_$i.json > gpurun_out/r2y_new_$i.json 2> gpurun_out/r2y_new_$i.err
    python - <<PY
import json
l=json.loads(open('gpurun_out/r2y_new_$i.json').read().strip().splitlines()[-1])
d=json.load(open('gpurun_out/r2y_detail_new_$i.json'))
k={x['name']:x['ms_per_step'] for x in d['kernels']}
print('new $i', round(l['ms_per_step'],1), l['clocks']['sm_mhz'], ' '.join(f"{n}={v:.2f}" for n,v in k.items() if v>0.3))
PY
done
