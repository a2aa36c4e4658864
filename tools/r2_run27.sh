#!/bin/bash
# round-2 evidence run: tests, smoke, bench records, ncu launch list / traffic / full capture
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest27.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke27.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke27.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "bench ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference.json
timeout 280 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02_lp_cfg3.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
timeout 200 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extra > gpurun_out/r2_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extra > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 200 python tools/run_plan.py --workload cfg3 --reps 2 > gpurun_out/r2_run_plan.log 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r02_traffic_cfg3.csv \
    python tools/run_plan.py --workload cfg3 --reps 2 > gpurun_out/r2_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"ssm_fused2|conv_tile|conv_pair" -s 30 -c 16 -o gpurun_out/r02_prof_cfg3 -f \
    python tools/run_plan.py --workload cfg3 --reps 2 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/*.ncu-rep
timeout 300 python tools/run_cfg4.py --textures 8 2>&1 | grep -E "^\{|Error" | tail -2 | cut -c1-500
