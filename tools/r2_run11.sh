#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 100 python tools/bench_tile.py
for e in 0 1 2 4 3 7; do ITG_TILE_DBG=1 ITG_TILE_EXP=$e timeout 100 python tools/bench_tile.py 2>&1 | grep -E "exp=|tile dbg" | cut -c1-420; done
# whole-step DRAM traffic of cfg3 (one pass, all launches): plain run first
timeout 200 python tools/run_plan.py --workload cfg3 --reps 2 > gpurun_out/r2_run_plan.log 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 32 -c 32 --csv --log-file gpurun_out/r02_traffic_cfg3.csv \
    python tools/run_plan.py --workload cfg3 --reps 2 > gpurun_out/r2_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
