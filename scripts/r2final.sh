python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2f_tests.log | tail -3
python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "bench rc=$?"; wc -c gpurun_out/r2f_bench_default.json
