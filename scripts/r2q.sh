python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "n2 rc=$?"; head -c 400 gpurun_out/r2q_bench_n2.json; echo
python -m pytest tests/test_gpu_e2e.py -m gpu -q -k "two_devices or sharding" > gpurun_out/r2q_twodev.log 2>&1; tail -2 gpurun_out/r2q_twodev.log
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py > gpurun_out/r2q_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/r2q_sanitizer.log
