#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_cuda_ops.py -x -q -m gpu -k "split or fp32" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_cuda_generator.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import bench
import infinite_texture_gans_b200 as itg
for wl in ("cfg2", "cfg3"):
    kw, th, tw, _ = bench.WORKLOADS[wl]
    if wl == "cfg3": th, tw = 15, 15
    sd = bench.make_weights(kw)
    outs = {}
    for prec in ("fp32-direct", "fp32", "fp16"):
        net = itg.ResidualPatchGenerator(**kw, precision=prec); net.load_state_dict(sd); net = net.cuda().eval()
        z, maps = bench.make_noise(kw, th, tw)
        with torch.no_grad():
            img = itg.utils.generate_full_grid(net, z, maps)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3): img = itg.utils.generate_full_grid(net, z, maps)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        outs[prec] = img.float().cpu()
        print(wl, th, tw, prec, f"{dt * 1e3:.2f} ms per call", flush=True)
    print(wl, "max|fp32 - fp32-direct| =", (outs["fp32"] - outs["fp32-direct"]).abs().max().item(), " max|fp16 - fp32-direct| =", (outs["fp16"] - outs["fp32-direct"]).abs().max().item(), flush=True)
PY
