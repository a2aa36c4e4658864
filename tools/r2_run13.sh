#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_cuda_ops.py -q -x > gpurun_out/r2_ops13.log 2>&1; echo "ops rc=$?"; tail -2 gpurun_out/r2_ops13.log
for lib in "" "build_variants/libitg_prev.so" "" "build_variants/libitg_prev.so"; do
  echo "== lib: ${lib:-new}"
  ITG_B200_LIB=$lib timeout 100 python tools/bench_tile.py 2>&1 | grep exp=
  for wl in cfg3 cfg5band; do
  ITG_B200_LIB=$lib timeout 280 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
  python - "$lib" $wl <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/r2_bench_tmp.json'))
    by = {k[:12]: (round(v['ms'], 2), round(v['frac'], 3)) for k, v in d['roofline']['by_kernel'].items()}
    print(sys.argv[1] or 'new    ', sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), by, 'e2e', round(d['e2e']['value']))
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'failed', e)
PY
  done
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest13.log
