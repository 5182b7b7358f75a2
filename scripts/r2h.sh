python -m pytest tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2h_tests.log
DC_LIB=$PWD/distilcodec_nabeel_b200/libdc_ng6.so python -m pytest tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/r2h_tests_ng6.log 2>&1; echo "tests ng6 rc=$?"; tail -3 gpurun_out/r2h_tests_ng6.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2h_detail_ng4.json > gpurun_out/r2h_bench_ng4.json 2> gpurun_out/r2h_bench_ng4.err; echo "bench ng4 rc=$?"
DC_LIB=$PWD/distilcodec_nabeel_b200/libdc_ng6.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2h_detail_ng6.json > gpurun_out/r2h_bench_ng6.json 2> gpurun_out/r2h_bench_ng6.err; echo "bench ng6 rc=$?"
python - <<'PY'
import json
for t in ('ng4','ng6'):
    d=json.load(open(f'gpurun_out/r2h_detail_{t}.json'))
    print(t, round(d['line']['ms_per_step'],1), [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels'] if 'pair' in k['name']])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"
CLIPS=64 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pairx --launch-skip 24 --launch-count 2 -f -o gpurun_out/r2_ncu_pairs python scripts/bench_stage.py generator pairs > gpurun_out/r2_ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
ls -la gpurun_out/*.ncu-rep
