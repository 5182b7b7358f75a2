python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2j_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2j_tests.log | tail -6
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2j_bench_reference_arm.json 2> gpurun_out/r2j_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 --detail-out gpurun_out/r2j_detail_recon_n1.json > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"; wc -c gpurun_out/r2j_bench_n1.json
python bench.py --workload vq_only --steps 100 --warmup 5 --detail-out gpurun_out/r2j_detail_vq_only.json > gpurun_out/r2j_bench_vq_only.json 2>> gpurun_out/r2j_side.err; echo "vq rc=$?"
python bench.py --workload wav2codes_30s --steps 8 --warmup 3 --detail-out gpurun_out/r2j_detail_wav2codes_30s.json > gpurun_out/r2j_bench_wav2codes_30s.json 2>> gpurun_out/r2j_side.err; echo "w2c rc=$?"
python bench.py --workload bulk_10min --steps 3 --warmup 3 --detail-out gpurun_out/r2j_detail_bulk_10min.json > gpurun_out/r2j_bench_bulk_10min.json 2>> gpurun_out/r2j_side.err; echo "bulk rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r2_launches.csv
