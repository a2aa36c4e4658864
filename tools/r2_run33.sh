#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_cuda_ops.py -x -q -m gpu -k "pair or grid_size" 2>&1 | tail -3
for rep in 1 2; do
for lib in build_variants/libitg_pair3x3only.so infinite_texture_gans_b200/libitg_b200.so; do
  echo "== $lib"; ITG_B200_LIB=$PWD/$lib timeout 100 python tools/bench_tile.py 2>&1 | grep "1x1"
  for wl in cfg2 cfg5band cfg3; do
  ITG_B200_LIB=$PWD/$lib timeout 280 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02_lp_$wl.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err || tail -5 gpurun_out/r2_bench_tmp.err
  python - "$wl" <<'PY'
import json, sys
wl = sys.argv[1]
d = json.load(open('gpurun_out/r2_bench_tmp.json')); lp = json.load(open(f'gpurun_out/r02_lp_{wl}.json'))['launches']
print(wl, 'ms/step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity', {}).get('max_abs'), ' '.join(f"{x['launch']}={x['ms']:.3f}" for x in lp if 'conv3' in x['launch']))
PY
  done
done
done
