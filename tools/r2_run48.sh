#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_cuda_ops.py -x -q -m gpu -k "pair or folded or grid_size or inside" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_cuda_generator.py tests/test_bench_shapes.py -x -q -m gpu 2>&1 | tail -4
for v in "ITG_NO_FOLD=1" "ITG_X=1" "ITG_NO_FOLD=1" "ITG_X=1"; do
  env $v timeout 280 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02_lp_cfg3.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err || tail -5 gpurun_out/r2_bench_tmp.err
  python - "$v" <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2_bench_tmp.json')); lp = json.load(open('gpurun_out/r02_lp_cfg3.json'))['launches']
print('cfg3', sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity', {}).get('max_abs'))
print('  ' + ' '.join(f"{x['launch']}={x['ms']:.3f}" for x in lp if x['kind'] in ('conv', 'att')))
PY
done
