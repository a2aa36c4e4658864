#!/usr/bin/env python
"""Time the thin-layer conv kernel on the shapes it runs in the SSM workload (developer tool; ITG_TILE_DBG=1 ITG_TILE_EXP=<bits> for experiments)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_texture_gans_b200 import _lib as L, packing as PK
from infinite_texture_gans_b200.ops import ConvOp, Grid, CudaBackend, c_store

be = CudaBackend()
dt = torch.float16
def grid(h, w, c):
    return Grid(torch.randn(h + 2, w + 2, c, device="cuda").to(dt), h, w, c)
def run(name, mode, H, W, cin, cout, res=None, reps=5):
    kin, kout = c_store(cin), c_store(cout)
    src = grid(H, W, kin)
    wt = torch.randn(cout, cin, 3, 3) / math.sqrt(9 * cin)
    w = (PK.pack_conv3x3(wt, dt) if mode == L.CONV3X3 else PK.pack_conv1x1(wt[:, :, :1, :1].contiguous(), dt)).cuda()
    op = ConvOp(mode=mode, src=src, w=w, k=kin, bias=torch.zeros(w.shape[1], device="cuda"), impl=L.IMPL_AUTO, name=name)
    op.out_h, op.out_w, op.out_c = H, W, kout
    op.out_raw = grid(H, W, kout)
    if res is not None:
        r = grid(H >> res, W >> res, kout)
        op.res_kind, op.res, op.res_shift, op.res_c, op.res_h, op.res_w = L.RES_GRID, r.buf, res, kout, H >> res, W >> res
    fn = be.compile_conv(op)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    taps = 9 if mode == L.CONV3X3 else 1
    print(f"{name:28s} exp={os.environ.get('ITG_TILE_EXP','0')} {ms:7.3f} ms  {2 * taps * cin * cout * H * W / ms / 1e9:7.1f} TFLOP/s", flush=True)
S = 3904
if os.environ.get("BENCH_TILE_ALL"):
    run("block3.conv2 3x3 104->104+res", L.CONV3X3, S // 4, S // 4, 104, 104, res=0)
    run("block4.conv1 3x3 104->52", L.CONV3X3, S // 2, S // 2, 104, 52)
    run("block4.conv2 3x3 52->52+res", L.CONV3X3, S // 2, S // 2, 52, 52, res=0)
run("block5.conv1 3x3 52->26", L.CONV3X3, S, S, 52, 26)
run("block5.conv3 1x1 52->26", L.CONV1X1, S, S, 52, 26)
run("block5.conv2 3x3 26->26+res", L.CONV3X3, S, S, 26, 26, res=0)
