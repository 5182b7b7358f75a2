DC_LIB=$PWD/distilcodec_nabeel_b200/libdc_cw32res.so python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "generator or decode or pair" > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2r_tests.log
for v in base cw32res base cw32res; do
  if [ $v = base ]; then unset DC_LIB; else export DC_LIB=$PWD/distilcodec_nabeel_b200/libdc_cw32res.so; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2r_detail_$v.json > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2r_detail_$v.json'))
print('$v', round(d['line']['ms_per_step'],1), d['line']['clocks'].get('sm_mhz'), [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels'] if 'tsw' in k['name']])
PY
done
