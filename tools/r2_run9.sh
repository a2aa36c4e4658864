#!/bin/bash
# two GPUs: the package's multi-GPU samplers and bench.py --gpus 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/band_check.py p2p cfg2 2>&1 | grep -E "band_check|Error|error" | tail -5
timeout 300 $TR --master-port 29534 tools/band_check.py p2p cfg3 2>&1 | grep -E "band_check|Error|error" | tail -5
timeout 300 $TR --master-port 29535 tools/band_check.py dist cfg2 2>&1 | grep -E "band_check|Error|error" | tail -5
timeout 300 python -m pytest tests/test_cuda_generator.py -q -k "two_gpu" 2>&1 | tail -3
timeout 600 $TR --master-port 29536 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_n2.json'))
    print('N=2 cfg3 bands: ms/step', round(d['ms_per_step'],3), 'MP/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e']['u8_value']), d['config']['workload'][-90:], 'launches', d['gpu_launches'])
    for x in d.get('extra',[]): print(x['name'], round(x['ms_per_step'],3), round(x['value']), 'e2e', round(x['e2e']['value']), 'u8', round(x['e2e']['u8_value']))
except Exception as e: print('parse failed', e)
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); print('N=1 cfg3: ms/step', round(d['ms_per_step'],3), 'MP/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3))"
timeout 600 $TR --master-port 29537 tools/run_cfg4.py --textures 16 2>&1 | grep -E "^\{|Error" | tail -2
timeout 300 python tools/run_cfg4.py --textures 8 2>&1 | grep -E "^\{|Error" | tail -2
