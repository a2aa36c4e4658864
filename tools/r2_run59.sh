#!/bin/bash
timeout 600 python -m pytest tests/test_cuda_ops.py -x -q -m gpu -k "pair or folded or inside" 2>&1 | tail -2
for rep in 1 2; do
for lib in build_variants/libitg_prebias.so infinite_texture_gans_b200/libitg_b200.so; do
  echo "== $lib"; ITG_B200_LIB=$PWD/$lib BENCH_TILE_ALL=1 timeout 100 python tools/bench_tile.py 2>&1 | grep exp= | grep -v "26->26"
  ITG_B200_LIB=$PWD/$lib timeout 280 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_tmp.json')); print('cfg3 ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))"
done
done
