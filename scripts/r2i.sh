DC_LIB=$PWD/distilcodec_nabeel_b200/libdc_cw32.so python -m pytest tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/r2i_tests_cw32.log 2>&1; echo "tests cw32 rc=$?"; tail -3 gpurun_out/r2i_tests_cw32.log
for v in base cw32 base cw32; do
  if [ $v = base ]; then unset DC_LIB; else export DC_LIB=$PWD/distilcodec_nabeel_b200/libdc_cw32.so; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2i_detail_$v.json > gpurun_out/r2i_bench_$v.json 2> gpurun_out/r2i_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2i_detail_$v.json'))
print('$v', round(d['line']['ms_per_step'],1), d['line']['clocks'].get('sm_mhz'), [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels'] if 'pair' in k['name']])
PY
done
