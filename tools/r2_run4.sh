#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for cg in 1 2; do for e in 0 1 2 3 4 8 16 7 31; do
ITG_SSM_CG=$cg ITG_TILE_DBG=1 ITG_SSM_EXP=$e timeout 120 python tools/bench_ssm.py 26,3904,3904,0 2>&1 | grep -E "^ssm|dbg" | tail -2
done; done > gpurun_out/r2_ssm_exp.log 2>&1
cat gpurun_out/r2_ssm_exp.log
