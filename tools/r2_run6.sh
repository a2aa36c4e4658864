#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ITG_SSM_CG=1 timeout 300 python -m pytest tests/test_cuda_ops.py -q -k fused_ssm > gpurun_out/r2_ssm_ops_cg1.log 2>&1; echo "ssm ops cg1 rc=$?"; tail -3 gpurun_out/r2_ssm_ops_cg1.log
timeout 300 python -m pytest tests/test_cuda_ops.py -q -x -k fused_ssm > gpurun_out/r2_ssm_ops_cg2.log 2>&1; rc2=$?; echo "ssm ops cg2 rc=$rc2"; tail -3 gpurun_out/r2_ssm_ops_cg2.log
for cg in 1 2; do
ITG_SSM_CG=$cg timeout 120 python tools/bench_ssm.py 2>&1 | grep -E "^ssm" 
ITG_SSM_CG=$cg ITG_TILE_DBG=1 timeout 120 python tools/bench_ssm.py 26,3904,3904,0 52,3904,3904,1 2>&1 | grep -E "dbg" | tail -2
done 2>&1 | tee gpurun_out/r2_ssm_exp3.log
if [ $rc2 -ne 0 ]; then export ITG_SSM_CG=1; echo "FALLING BACK TO CG1 for the rest"; fi
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_bench_shapes.py > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest6.log
timeout 120 python tools/bench_conv.py > gpurun_out/r2_bench_conv.log 2>&1; cat gpurun_out/r2_bench_conv.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r2_lp_cfg3.json > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_default.json'))
    def show(e,name):
        print(name, 'ms/step', round(e['ms_per_step'],3), 'MP/s', round(e['value']), 'e2e', round(e['e2e']['value']), 'u8', round(e['e2e']['u8_value']), 'roof', round(e['roofline']['frac'],3), 'step frac', round(e['roofline']['step_frac_of_peak'],3), 'parity', e['parity'])
    show(d,'cfg3')
    print({k:(round(v['frac'],3), round(v['ms'],2)) for k,v in d['roofline']['by_kernel'].items()})
    for x in d.get('extra',[]): show(x,x['name'])
    print('cpu', d.get('cpu_baseline'))
except Exception as ex: print('parse failed', ex)
PY
