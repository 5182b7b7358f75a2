python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2a_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"; wc -c gpurun_out/r2a_bench_n1.json
cp profiles/bench_detail_recon_n1.json gpurun_out/r2a_detail_recon_n1.json
for w in vq_only wav2codes_30s bulk_10min; do
  python bench.py --workload $w --steps 4 --warmup 3 --detail-out gpurun_out/r2a_detail_$w.json > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err; echo "$w rc=$?"; tail -c 600 gpurun_out/r2a_bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"; nproc; head -c 400 gpurun_out/r2a_ref.json
