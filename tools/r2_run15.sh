#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for v in "" "ITG_TILE_MINRING=3" "ITG_TILE_MINRING=4" "ITG_TILE_SMEM_KB=200" ""; do
  echo "== ${v:-default}"; env $v timeout 100 python tools/bench_tile.py 2>&1 | grep exp=
done
ITG_TILE_DBG=1 timeout 100 python tools/bench_tile.py 2>&1 | grep -E "tile dbg" | awk 'NR%7==1' | cut -c1-420
timeout 280 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_lp_cfg3.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_tmp.json')); lp = json.load(open('gpurun_out/r02_lp_cfg3.json'))['launches']
print('cfg3 ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3))
for x in lp:
    if x['kind'] in ('conv', 'att'): print(f"  {x['launch']:18s} {x['ms']:.3f}")
PY
