#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ITG_SSM_CG=1 timeout 300 python -m pytest tests/test_cuda_ops.py -q -k fused_ssm > gpurun_out/r2_ssm_ops_cg1.log 2>&1; echo "ssm ops cg1 rc=$?"; tail -3 gpurun_out/r2_ssm_ops_cg1.log
timeout 300 python -m pytest tests/test_cuda_ops.py -q -x -k fused_ssm > gpurun_out/r2_ssm_ops_cg2.log 2>&1; rc2=$?; echo "ssm ops cg2 rc=$rc2"; tail -3 gpurun_out/r2_ssm_ops_cg2.log
if [ $rc2 -ne 0 ]; then export ITG_SSM_CG=1; echo "FALLING BACK TO CG1 for the rest"; fi
ITG_TILE_DBG=1 timeout 200 python tools/run_plan.py --workload cfg3 --reps 1 > gpurun_out/r2_cfg3_dbg.log 2>&1; grep "ssm dbg" gpurun_out/r2_cfg3_dbg.log | tail -4
ITG_SSM_CG=1 ITG_TILE_DBG=1 timeout 200 python tools/run_plan.py --workload cfg3 --reps 1 > gpurun_out/r2_cfg3_dbg_cg1.log 2>&1; grep "ssm dbg" gpurun_out/r2_cfg3_dbg_cg1.log | tail -4
timeout 280 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_lp_cfg3_fused.json > gpurun_out/r2_bench_cfg3_fused.json 2> gpurun_out/r2_bench_cfg3_fused.err; cat gpurun_out/r2_bench_cfg3_fused.json | head -c 330; echo
ITG_SSM_CG=1 timeout 280 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_lp_cfg3_cg1.json > gpurun_out/r2_bench_cfg3_cg1.json 2> gpurun_out/r2_bench_cfg3_cg1.err; cat gpurun_out/r2_bench_cfg3_cg1.json | head -c 330; echo
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_bench_shapes.py > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest3.log
