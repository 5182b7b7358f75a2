python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2f_tests.log | tail -3
python bench.py --steps 20 --warmup 5 --detail-out gpurun_out/r2f_detail_recon_n1.json > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"; wc -c gpurun_out/r2f_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"
