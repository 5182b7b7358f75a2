python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2o_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" gpurun_out/r2o_tests.log | tail -4
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2o_bench_reference_arm.json 2> gpurun_out/r2o_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 --detail-out gpurun_out/r2o_detail_recon_n1.json > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo "bench rc=$?"; wc -c gpurun_out/r2o_bench_n1.json
python bench.py --workload vq_only --steps 100 --warmup 5 --detail-out gpurun_out/r2o_detail_vq_only.json > gpurun_out/r2o_bench_vq_only.json 2>> gpurun_out/r2o_side.err; echo "vq rc=$?"
python bench.py --workload wav2codes_30s --steps 8 --warmup 3 --detail-out gpurun_out/r2o_detail_wav2codes_30s.json > gpurun_out/r2o_bench_wav2codes_30s.json 2>> gpurun_out/r2o_side.err; echo "w2c rc=$?"
python bench.py --workload bulk_10min --steps 3 --warmup 3 --detail-out gpurun_out/r2o_detail_bulk_10min.json > gpurun_out/r2o_bench_bulk_10min.json 2>> gpurun_out/r2o_side.err; echo "bulk rc=$?"
python bench.py --mode fp32 --clips 32 --steps 3 --warmup 3 --no-cpu-baseline --detail-out gpurun_out/r2o_detail_fp32.json > gpurun_out/r2o_bench_fp32.json 2>> gpurun_out/r2o_side.err; echo "fp32 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"
for cl in 1 2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:vq_score --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_ncu_vq_cl$cl python bench.py --workload vq_only --steps 1 --warmup 3 --no-cpu-baseline --opt cta_pairs=$cl --detail-out gpurun_out/tmp_detail.json > gpurun_out/r2_ncu_vq_cl$cl.log 2>&1; echo "ncu vq cl=$cl rc=$?"
done
CLIPS=64 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tsw --launch-skip 120 --launch-count 2 -f -o gpurun_out/r2_ncu_tsw python scripts/bench_stage.py generator tsw > gpurun_out/r2_ncu_tsw.log 2>&1; echo "ncu tsw rc=$?"
ls -la gpurun_out/*.ncu-rep
