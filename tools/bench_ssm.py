#!/usr/bin/env python
"""Time single itg_ssm_fwd launches (developer tool).  ITG_TILE_DBG=1 prints the per-role cycle counters; with it, ITG_SSM_EXP=<bits>
switches parts of the kernel off (wrong results, timing only) to see what bounds it."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_texture_gans_b200 import _lib as L, packing as PK
from infinite_texture_gans_b200.ops import CudaBackend, Grid, SsmOp, c_store

be = CudaBackend()
dt = torch.float16
shapes = [(26, 3904, 3904, 0), (52, 3904, 3904, 1), (104, 1952, 1952, 1)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for C, H, W, shift in shapes:
    cs = c_store(C)
    g = torch.Generator().manual_seed(1)
    w_mlp = PK.pack_ssm_mlp(torch.randn(128, 1, 3, 3, generator=g) / 3.0, 0.2 * torch.randn(128, generator=g), dt).cuda()
    w_emb, b_emb = PK.pack_ssm_embed(torch.randn(2 * C, 128, 3, 3, generator=g) / math.sqrt(9 * 128), 0.1 * torch.randn(2 * C, generator=g), dt)
    xh, xw = (H + shift) >> shift, (W + shift) >> shift
    x = Grid(torch.randn(xh + 2, xw + 2, cs, device="cuda").to(dt), xh, xw, cs)
    op = SsmOp(map=torch.randn(H + 4, W + 4, device="cuda"), w_mlp=w_mlp, w_embed=w_emb.cuda(), b_embed=b_emb.cuda(), x=x, x_shift=shift,
               mean=torch.zeros(cs, device="cuda"), rstd=torch.ones(cs, device="cuda"),
               out=Grid(torch.empty(H + 2, W + 2, cs, device="cuda", dtype=dt), H, W, cs), leak=0.02, border=L.BORDER_REPLICATE)
    fn = be.compile_ssm(op)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * H * W * 1152 * 2 * C + 2.0 * (H + 2) * (W + 2) * 9 * 128
    print(f"ssm C={C} {H}x{W} shift={shift} cg={os.environ.get('ITG_SSM_CG', '2')} exp={os.environ.get('ITG_SSM_EXP', '0')}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s algorithmic", flush=True)
