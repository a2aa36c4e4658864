#!/bin/bash
for rep in 1 2 3; do
for lib in infinite_texture_gans_b200/libitg_b200.so build_variants/libitg_prebias.so; do
  ITG_B200_LIB=$PWD/$lib timeout 280 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r2_bench_tmp.json')); r=d['roofline']; print(sys.argv[1][-18:], 'cfg3 ms/step', round(d['ms_per_step'],3), 'conv_ms', round(r['conv_ms_per_step'],3), {k[:8]: round(v['ms'],3) for k,v in r['by_kernel'].items()}, d['clocks']['sm_mhz'])" $lib
done
done
