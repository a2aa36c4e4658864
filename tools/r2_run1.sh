#!/bin/bash
# Round 2, GPU call 1: fused SSM kernel first (own process: a trap poisons the CUDA context), then the whole gpu suite, the MMA probe,
# and the SSM workload with the fused and the two-launch path.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_cuda_ops.py -q -k fused_ssm > gpurun_out/r2_ssm_ops.log 2>&1; echo "ssm ops rc=$?"; tail -15 gpurun_out/r2_ssm_ops.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest1.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/umma_probe tools/umma_probe.cu -lcuda && timeout 120 /tmp/umma_probe > gpurun_out/r2_umma_probe.txt 2>&1; tail -5 gpurun_out/r2_umma_probe.txt
ITG_TILE_DBG=1 timeout 200 python tools/run_plan.py --workload cfg3 --reps 1 > gpurun_out/r2_cfg3_dbg.log 2>&1; grep "ssm dbg" gpurun_out/r2_cfg3_dbg.log | tail -30
timeout 280 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_lp_cfg3_fused.json > gpurun_out/r2_bench_cfg3_fused.json 2> gpurun_out/r2_bench_cfg3_fused.err; cat gpurun_out/r2_bench_cfg3_fused.json | head -c 600; echo
ITG_SSM_UNFUSED=1 timeout 280 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_lp_cfg3_unfused.json > gpurun_out/r2_bench_cfg3_unfused.json 2> gpurun_out/r2_bench_cfg3_unfused.err; cat gpurun_out/r2_bench_cfg3_unfused.json | head -c 600; echo
