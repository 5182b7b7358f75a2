python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2e_tests.log | tail -12
for m in 0 1 2; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt pairx=$m --detail-out gpurun_out/r2e_detail_pairx$m.json > gpurun_out/r2e_bench_pairx$m.json 2> gpurun_out/r2e_bench_pairx$m.err; echo "bench pairx=$m rc=$?"; tail -2 gpurun_out/r2e_bench_pairx$m.err
done
python scripts/audio_loader_bench.py 64 30 > gpurun_out/r2e_audio_loader.json 2>&1; cat gpurun_out/r2e_audio_loader.json
