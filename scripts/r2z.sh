timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
  for cp in 1 2; do
    timeout 600 python bench.py --mode fp32 --clips 64 --steps 4 --warmup 3 --no-cpu-baseline --opt cta_pairs=$cp --detail-out gpurun_out/r2z_detail_cp${cp}_$i.json > gpurun_out/r2z_cp${cp}_$i.json 2> gpurun_out/r2z_cp${cp}_$i.err
    python - <<PY
import json
l=json.loads(open('gpurun_out/r2z_cp${cp}_$i.json').read().strip().splitlines()[-1])
d=json.load(open('gpurun_out/r2z_detail_cp${cp}_$i.json'))
k={x['name']:x['ms_per_step'] for x in d['kernels']}
print('cp$cp $i', round(l['ms_per_step'],1), round(l['value'],1), l['clocks']['sm_mhz'], l.get('codes_checksum'), ' '.join(f"{n}={v:.1f}" for n,v in k.items() if v>5))
PY
  done
done
