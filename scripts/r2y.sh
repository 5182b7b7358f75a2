python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_config_size.py tests/test_gpu_modules.py tests/test_gpu_dropin_joined.py tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -6
for i in 1 2; do
  for v in 0 1; do
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt branch_bf16=$v --detail-out gpurun_out/r2y_detail_bb${v}_$i.json > gpurun_out/r2y_bb${v}_$i.json 2> gpurun_out/r2y_bb${v}_$i.err
    python - <<PY
import json
l=json.loads(open('gpurun_out/r2y_bb${v}_$i.json').read().strip().splitlines()[-1])
d=json.load(open('gpurun_out/r2y_detail_bb${v}_$i.json'))
ly={x['name']:x['ms_per_step'] for x in d['layers']}
print('bb$v $i', round(l['ms_per_step'],1), l['clocks']['sm_mhz'], l.get('codes_checksum'), round(sum(v for n,v in ly.items() if ' e44' in n or ' e20' in n or 'e12]' in n or 'e36]' in n or 'e40]' in n),2), ' '.join(f"{n}={v:.2f}" for n,v in ly.items() if ('e44' in n or 'e12]' in n or 'e40]' in n)))
PY
  done
done
