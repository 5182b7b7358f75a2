timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 --detail-out gpurun_out/r2_bench_detail_recon_n8.json > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "n8 rc=$?"
tail -1 gpurun_out/r2_bench_n8.json | cut -c1-1500
