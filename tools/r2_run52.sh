#!/bin/bash
# per-role counters of the pair conv launches inside a cfg3 pass (debug build)
for e in 0 1 2; do
echo "== exp $e"
ITG_TILE_EXP=$e ITG_TILE_DBG=1 ITG_B200_LIB=$PWD/build_variants/libitg_dbg.so timeout 200 python tools/run_plan.py --workload cfg3 --reps 1 2>&1 | grep "pair dbg" | tail -7 | sed 's/nblocks.*kcycles CTA0://' | cut -c15-330
done
