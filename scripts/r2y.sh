for lib in base new; do
  if [ $lib = new ]; then unset DC_LIB; else export DC_LIB=$PWD/ab/libdc_$lib.so; fi
  echo "== $lib"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
done
unset DC_LIB
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_config_size.py tests/test_gpu_ops.py tests/test_gpu_dropin_joined.py -m gpu -q -s 2>&1 | grep -E "passed|failed|W1 encoder"
for i in 1 2; do
  for lib in base new; do
    if [ $lib = new ]; then unset DC_LIB; else export DC_LIB=$PWD/ab/libdc_$lib.so; fi
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2y_detail_${lib}_$i.json > gpurun_out/r2y_${lib}_$i.json 2> gpurun_out/r2y_${lib}_$i.err
    python - <<PY
import json
l=json.loads(open('gpurun_out/r2y_${lib}_$i.json').read().strip().splitlines()[-1])
d=json.load(open('gpurun_out/r2y_detail_${lib}_$i.json'))
k={x['name']:x['ms_per_step'] for x in d['kernels']}
print('$lib $i', round(l['ms_per_step'],1), l['clocks']['sm_mhz'], l.get('codes_checksum'), ' '.join(f"{n}={v:.2f}" for n,v in k.items() if n.startswith('gemm_tc')))
PY
  done
done
