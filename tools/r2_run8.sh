#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ITG_SSM_CG=1 timeout 300 python -m pytest tests/test_cuda_ops.py -q -k fused_ssm > gpurun_out/r2_ssm_ops_cg1.log 2>&1; echo "ssm ops cg1 rc=$?"; tail -2 gpurun_out/r2_ssm_ops_cg1.log
timeout 300 python -m pytest tests/test_cuda_ops.py -q -x -k fused_ssm > gpurun_out/r2_ssm_ops_cg2.log 2>&1; rc2=$?; echo "ssm ops cg2 rc=$rc2"; tail -2 gpurun_out/r2_ssm_ops_cg2.log
if [ $rc2 -ne 0 ]; then export ITG_SSM_CG=1; echo "FALLING BACK TO CG1 for the rest"; fi
timeout 120 python tools/bench_ssm.py 2>&1 | grep -E "^ssm"
ITG_SSM_NGRAN=16 timeout 120 python tools/bench_ssm.py 2>&1 | grep -E "^ssm" | sed 's/^/ngran16 /'
timeout 120 python tools/bench_ssm.py 2>&1 | grep -E "^ssm"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_bench_shapes.py > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest8.log
for v in "" "ITG_TILE_MINRING=3" "ITG_TILE_MINRING=4" "ITG_TILE_PIPES=1"; do
  env $v timeout 280 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_lp_tmp.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/r2_bench_tmp.json')); lp = json.load(open('gpurun_out/r2_lp_tmp.json'))['launches']
    conv = {x['launch']: round(x['ms'], 3) for x in lp if x['kind'] == 'conv' and x['ms'] > 0.3}
    print(sys.argv[1] or 'default', 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'ssm', round(sum(x['ms'] for x in lp if x['kind'] == 'ssm'), 2), 'conv', round(sum(x['ms'] for x in lp if x['kind'] == 'conv'), 2), conv)
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done
