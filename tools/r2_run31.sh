#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29533 tools/band_check.py p2p cfg3 2>&1 | grep -E "band_check|Error" | tail -5
timeout 300 $TR --nproc-per-node 2 --master-port 29534 tools/band_check.py p2p cfg2 2>&1 | grep -E "band_check|Error" | tail -5
timeout 300 $TR --nproc-per-node 2 --master-port 29535 tools/band_check.py dist cfg3 2>&1 | grep -E "band_check|Error" | tail -5
timeout 600 $TR --nproc-per-node 2 --master-port 29560 tools/run_cfg5.py --rows 40 --cols 129 > gpurun_out/r02_cfg5_n2.log 2>&1; echo "cfg5 rc=$?"; grep -E "^\{" gpurun_out/r02_cfg5_n2.log | tail -1 | cut -c1-900
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest31.log
