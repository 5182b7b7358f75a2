for i in 1 2; do
  for pf in 1 2 0; do
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt epi_prefetch=$pf --detail-out gpurun_out/r2y_detail_pf${pf}_$i.json > gpurun_out/r2y_pf${pf}_$i.json 2> gpurun_out/r2y_pf${pf}_$i.err
    python - <<PY
import json
l=json.loads(open('gpurun_out/r2y_pf${pf}_$i.json').read().strip().splitlines()[-1])
d=json.load(open('gpurun_out/r2y_detail_pf${pf}_$i.json'))
k={x['name']:x['ms_per_step'] for x in d['kernels']}
ly={x['name']:x['ms_per_step'] for x in d['layers']}
print('pf$pf $i', round(l['ms_per_step'],1), l['clocks']['sm_mhz'], ' '.join(f"{n}={v:.2f}" for n,v in k.items() if v>20), '|', ' '.join(f"{n[14:]}={v:.2f}" for n,v in ly.items() if n.startswith('conv_tsw<2,32,3>')))
PY
  done
done
