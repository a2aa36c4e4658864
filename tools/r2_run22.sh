#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_cuda_ops.py -x -q -m gpu -k "pair" 2>&1 | tail -3
for v in "ITG_CONV_PAIR=0" "ITG_CONV_PAIR=2" "ITG_CONV_PAIR=2 ITG_PAIR_INFLIGHT=1"; do
  echo "== $v"; env $v BENCH_TILE_ALL=1 timeout 100 python tools/bench_tile.py 2>&1 | grep exp=
done
for e in 0 1 2 3; do
echo "== exp $e"; env ITG_TILE_EXP=$e ITG_CONV_PAIR=2 ITG_TILE_DBG=1 ITG_B200_LIB=$PWD/build_variants/libitg_dbg.so BENCH_TILE_ALL=1 timeout 100 python tools/bench_tile.py 2>&1 | grep -E "pair dbg" | awk 'NR%7==1' | sed 's/nblocks.*kcycles CTA0://' | cut -c15-400
done
