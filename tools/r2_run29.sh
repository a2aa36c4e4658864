#!/bin/bash
# eight GPUs: band parity, bench.py --gpus 8 and 1 on the same box, BASELINE config 5 at full size
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29533 tools/band_check.py p2p cfg3 2>&1 | grep -E "band_check|Error|error" | tail -5
for n in 8; do
timeout 900 $TR --nproc-per-node $n --master-port $((29540 + n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo "bench n$n rc=$?"; tail -2 gpurun_out/r02_bench_n$n.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_samebox.json 2> gpurun_out/r02_bench_n1_samebox.err
python - <<'PY'
import json
for n in (1, 8):
    try:
        d = json.load(open(f'gpurun_out/r02_bench_n{n}.json' if n > 1 else 'gpurun_out/r02_bench_n1_samebox.json'))
        print(f'N={n} cfg3 bands: ms/step', round(d['ms_per_step'], 3), 'MP/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e']['u8_value']), 'launches', d['gpu_launches'])
        for x in d.get('extra', []): print('   ', x['name'], round(x['ms_per_step'], 3), round(x['value']), 'e2e', round(x['e2e']['value']))
    except Exception as e: print(n, 'parse failed', e)
PY
timeout 900 $TR --nproc-per-node 8 --master-port 29560 tools/run_cfg5.py > gpurun_out/r02_cfg5_n8.log 2>&1; echo "cfg5 rc=$?"; grep -E "^\{" gpurun_out/r02_cfg5_n8.log | tail -1 | cut -c1-900
