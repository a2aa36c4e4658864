#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for rep in 1 2; do
for lib in build_variants/libitg_nopad.so infinite_texture_gans_b200/libitg_b200.so; do
  echo "== $lib"; ITG_B200_LIB=$PWD/$lib timeout 100 python tools/bench_tile.py 2>&1 | grep exp=
  for wl in cfg2 cfg5band cfg3; do
  ITG_B200_LIB=$PWD/$lib timeout 280 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err || tail -5 gpurun_out/r2_bench_tmp.err
  python - "$wl" <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2_bench_tmp.json'))
print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity', {}).get('max_abs'))
PY
  done
done
done
