#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
# N granularity 16 for the pair MMA (own process: an illegal instruction would poison the context)
ITG_SSM_NGRAN=16 timeout 300 python -m pytest tests/test_cuda_ops.py -q -x -k fused_ssm > gpurun_out/r2_ssm_ops_n16.log 2>&1; echo "ssm ops ngran16 rc=$?"; tail -3 gpurun_out/r2_ssm_ops_n16.log
ITG_SSM_NGRAN=16 timeout 120 python tools/bench_ssm.py 2>&1 | grep -E "^ssm"
ITG_TILE_DBG=1 timeout 120 python tools/bench_ssm.py 26,3904,3904,0 2>&1 | grep -E "dbg" | tail -1
ITG_TILE_DBG=1 timeout 200 python tools/run_plan.py --workload cfg3 --reps 1 > gpurun_out/r2_cfg3_dbg.log 2>&1; grep "tile dbg" gpurun_out/r2_cfg3_dbg.log | tail -8
# profiles: launch list, then one full capture of the SSM pair kernel and the thin-layer conv kernel (plain runs first)
timeout 200 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 200 python tools/run_plan.py --workload cfg3 --reps 2 > gpurun_out/r2_run_plan.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"ssm_fused2|conv_tile" -s 30 -c 16 -o gpurun_out/r02_prof_cfg3 -f \
    python tools/run_plan.py --workload cfg3 --reps 2 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/*.ncu-rep
