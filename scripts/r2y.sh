timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
  for lib in base new; do
    if [ $lib = new ]; then unset DC_LIB; else export DC_LIB=$PWD/ab/libdc_$lib.so; fi
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2y_detail_${lib}_$i.json > gpurun_out/r2y_${lib}_$i.json 2> gpurun_out/r2y_${lib}_$i.err
    python - <<PY
import json
l=json.loads(open('gpurun_out/r2y_${lib}_$i.json').read().strip().splitlines()[-1])
d=json.load(open('gpurun_out/r2y_detail_${lib}_$i.json'))
k={x['name']:x['ms_per_step'] for x in d['kernels']}
ly={x['name']:x['ms_per_step'] for x in d['layers']}
print('$lib $i', round(l['ms_per_step'],1), l['clocks']['sm_mhz'], l.get('codes_checksum'), 'tsw216=%.2f'%k.get('conv_tsw<2,16,4>x2',0), ' '.join(f"{n[18:]}={v:.2f}" for n,v in ly.items() if n.startswith('conv_tsw<2,16,4>')))
PY
  done
done
