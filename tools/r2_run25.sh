#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for wl in cfg3 cfg2 cfg5band; do
for v in "ITG_CONV_PAIR=0" "ITG_CONV_PAIR=1" "ITG_CONV_PAIR=1 ITG_PAIR_MIN_TILES=296" "ITG_CONV_PAIR=2"; do
  env $v timeout 280 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02_lp_$wl.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err || tail -5 gpurun_out/r2_bench_tmp.err
  python - "$wl" "$v" <<'PY'
import json, sys
wl = sys.argv[1]
d = json.load(open('gpurun_out/r2_bench_tmp.json')); lp = json.load(open(f'gpurun_out/r02_lp_{wl}.json'))['launches']
print(wl, sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity', {}).get('max_abs'))
if wl == 'cfg3' or 'PAIR=0' in sys.argv[2] or 'PAIR=2' in sys.argv[2]:
    print('  ' + ' '.join(f"{x['launch']}={x['ms']:.3f}" for x in lp if x['kind'] in ('conv', 'att')))
PY
done
done
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_bench_shapes.py 2>&1 | tail -3
