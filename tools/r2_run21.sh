#!/bin/bash
for e in 0 1 2 3 4 7; do
echo "== exp $e"; env ITG_TILE_EXP=$e ITG_CONV_PAIR=2 ITG_TILE_DBG=1 ITG_B200_LIB=$PWD/build_variants/libitg_dbg.so BENCH_TILE_ALL=1 timeout 100 python tools/bench_tile.py 2>&1 | grep -E "pair dbg" | awk 'NR%7==1' | cut -c15-40,168-400
done
