#!/usr/bin/env python
"""Time single conv launches (developer tool): thin-layer shapes of the 241 Generator with different epilogues."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_texture_gans_b200 import _lib as L, packing as PK
from infinite_texture_gans_b200.ops import ConvOp, Grid, CudaBackend, c_store

be = CudaBackend()
dt = torch.float16

def grid(h, w, c):
    return Grid(torch.randn(h + 2, w + 2, c, device="cuda").to(dt), h, w, c)

def run(name, mode, H, W, cin, cout, raw=False, act=False, res=None, img=False, impl=L.IMPL_AUTO, reps=20):
    kin, kout = c_store(cin), c_store(cout)
    src = grid(H, W, kin)
    wt = torch.randn(cout, cin, 3, 3) / math.sqrt(9 * cin)
    if mode == L.CONV3X3: w, s = PK.pack_conv3x3(wt, dt), 1
    elif mode == L.CONV1X1: w, s = PK.pack_conv1x1(wt[:, :, :1, :1].contiguous(), dt), 1
    else: w, s = PK.pack_upconv(wt, dt), 2
    w = w.cuda()
    op = ConvOp(mode=mode, src=src, w=w, k=kin, bias=torch.zeros(w.shape[1], device="cuda"), impl=impl, name=name)
    oh, ow = s * H, s * W
    op.out_h, op.out_w, op.out_c = oh, ow, kout
    if img:
        op.out_img, op.img_c, op.img_layout = torch.empty(1, cout, oh, ow, device="cuda"), cout, L.IMG_MERGED
    if raw: op.out_raw = grid(oh, ow, kout)
    if act:
        op.out_act = grid(oh, ow, kout)
        op.scale, op.shift, op.leak = torch.ones(w.shape[1], device="cuda"), torch.zeros(w.shape[1], device="cuda"), 0.02
        op.border = L.BORDER_REPLICATE
    if res is not None:
        rh, rw = oh >> res, ow >> res
        r = grid(rh, rw, kout)
        op.res_kind, op.res, op.res_shift, op.res_c, op.res_h, op.res_w = L.RES_GRID, r.buf, res, kout, rh, rw
    fn = be.compile_conv(op)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    px = oh * ow
    byts = (H * W * kin + (px * kout if raw else 0) + (px * kout if act else 0) + ((px >> (2 * res)) * kout if res is not None else 0)) * 2 + (px * cout * 4 if img else 0)
    print(f"{name:34s} {us:8.1f} us   {byts / us / 1e3:7.1f} GB/s   {2 * 9 * cin * cout * px / us / 1e6:7.1f} TFLOP/s(alg, 3x3-equivalent)")

H, W = 896, 2688
run("3x3 16->16 ACT", L.CONV3X3, H, W, 13, 13, act=True)
run("3x3 16->16 ACT|RES(shift1)", L.CONV3X3, H, W, 13, 13, act=True, res=1)
run("3x3 16->16 ACT|RES(shift0)", L.CONV3X3, H, W, 13, 13, act=True, res=0)
run("3x3 16->16 RAW", L.CONV3X3, H, W, 13, 13, raw=True)
run("3x3 16->16 RAW|ACT|RES", L.CONV3X3, H, W, 13, 13, raw=True, act=True, res=1)
run("3x3 16->3 IMG", L.CONV3X3, H, W, 13, 3, img=True)
run("3x3 16->16 ACT (streaming kernel)", L.CONV3X3, H, W, 13, 13, act=True, impl=L.IMPL_UMMA)
run("up 32->16 ACT", L.UPCONV, H // 2, W // 2, 26, 13, act=True)
run("3x3 32->32 RAW|ACT|RES", L.CONV3X3, H // 2, W // 2, 26, 26, raw=True, act=True, res=1)
run("up 64->32 ACT", L.UPCONV, H // 4, W // 4, 52, 26, act=True)
run("3x3 64->64 RAW|ACT|RES", L.CONV3X3, H // 4, W // 4, 52, 52, raw=True, act=True, res=1)
run("1x1 32->16 RAW", L.CONV1X1, H // 2, W // 2, 26, 13, raw=True)
