#!/bin/bash
# GPU side of the per-round profile refresh (run under gpurun); tools/make_profile_summary.sh turns the outputs into profiles/.
# Every ncu pass runs only after the same command has exited 0 without ncu.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --profile-out gpurun_out/launch_profile_cfg2.json > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
timeout 200 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
if [ -z "$SKIP_FULL" ]; then
timeout 200 python tools/run_plan.py > gpurun_out/run_plan.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tile|conv_umma|attention_mma" -s 20 -c 20 \
    -o gpurun_out/r01_prof_cfg2 -f python tools/run_plan.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ls -la gpurun_out | tail -12
