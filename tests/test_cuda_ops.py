"""Launch-by-launch parity of the CUDA kernels (through the C ABI) with the CPU emulation of the same
descriptors (tests/emulator.py), on identical operands.

Tolerances: fp32 direct conv <= 2e-5 relative to the output scale (fp32 re-association only); 16-bit paths
are fed the *same* 16-bit-rounded operands as the emulator, so what remains is fp32 accumulation order plus
one rounding of the stored result: <= 2^-9 (fp16) / 2^-6 (bf16) of the output scale.
"""
import math

import pytest
import torch

from emulator import EmulatorBackend
from infinite_texture_gans_b200 import _lib as L
from infinite_texture_gans_b200 import packing as PK
from infinite_texture_gans_b200.ops import AttentionOp, ConvOp, Grid, SsmOp, c_store, k_pad_of

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
REL = {torch.float32: 2e-5, torch.float16: 2.0 ** -9, torch.bfloat16: 2.0 ** -6}


@pytest.fixture(scope="module")
def be():
    from infinite_texture_gans_b200.ops import CudaBackend
    return CudaBackend()


def _grid(h, w, c, dtype, gen, device="cpu", scale=1.0):
    buf = (torch.randn((h + 2, w + 2, c), generator=gen) * scale).to(dtype)
    return Grid(buf.to(device), h, w, c)


def _clone_grid(g: Grid, device):
    return Grid(g.buf.clone().to(device), g.h, g.w, g.c)


def _empty_like(g: Grid, device, fill=7.0):
    return Grid(torch.full_like(g.buf, fill).to(device), g.h, g.w, g.c)


def _cmp(name, got: torch.Tensor, ref: torch.Tensor, dtype):
    got, ref = got.float().cpu(), ref.float()
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item()
    assert err <= REL[dtype] * scale + 1e-6, f"{name}: max-abs {err:.3e} at output scale {scale:.3e}"


def _run_both(be, op_cpu: ConvOp, to_dev):
    """Run op on the emulator (CPU tensors) and its device copy through the C ABI."""
    EmulatorBackend().conv(op_cpu)
    op_gpu = to_dev(op_cpu)
    be.conv(op_gpu)
    torch.cuda.synchronize()
    return op_gpu


CONV_CASES = [
    # (mode, H, W, cin, cout, extras)
    ("3x3", 12, 20, 128, 416, dict(raw=True, act=True, border=L.BORDER_REPLICATE)),
    ("3x3", 12, 20, 416, 416, dict(act=True, border=L.BORDER_CONSTANT)),
    ("3x3", 17, 33, 64, 64, dict(raw=True, act=True, res=0, border=L.BORDER_REPLICATE)),
    ("3x3", 16, 24, 104, 104, dict(raw=True, res=1, border=L.BORDER_NONE)),
    ("3x3", 33, 70, 52, 52, dict(act=True, res=1, border=L.BORDER_REPLICATE)),
    ("3x3", 40, 48, 26, 26, dict(raw=True, act=True, border=L.BORDER_REPLICATE)),
    ("3x3", 64, 40, 13, 13, dict(act=True, border=L.BORDER_REPLICATE)),
    ("3x3", 9, 5, 8, 8, dict(raw=True, act=True, border=L.BORDER_REPLICATE)),
    ("3x3", 48, 80, 13, 3, dict(img=L.IMG_MERGED)),
    ("3x3", 32, 64, 16, 3, dict(img=L.IMG_PATCHES, patch=16)),
    ("1x1", 12, 20, 416, 208, dict(raw=True)),
    ("1x1", 19, 21, 26, 13, dict(raw=True)),
    ("1x1", 30, 14, 16, 128, dict(act=True, relu=True)),
    ("up", 12, 20, 416, 208, dict(act=True, border=L.BORDER_REPLICATE)),
    ("up", 9, 13, 104, 52, dict(act=True, border=L.BORDER_CONSTANT)),
    ("up", 24, 40, 26, 13, dict(act=True, border=L.BORDER_REPLICATE)),
    ("up", 7, 6, 8, 8, dict(act=True, raw=True, border=L.BORDER_REPLICATE)),
    # thin layers of the 241 Generator at sizes with many / partial 16x8 tiles (halo-tile kernel)
    ("3x3", 130, 300, 52, 52, dict(raw=True, act=True, res=1, border=L.BORDER_REPLICATE)),
    ("3x3", 129, 67, 13, 13, dict(act=True, res=1, border=L.BORDER_CONSTANT)),
    ("3x3", 200, 264, 26, 26, dict(raw=True, act=True, border=L.BORDER_REPLICATE)),
    ("up", 65, 131, 52, 26, dict(act=True, border=L.BORDER_REPLICATE)),
    ("up", 130, 50, 26, 13, dict(act=True, border=L.BORDER_CONSTANT)),
    ("1x1", 70, 90, 52, 26, dict(raw=True)),
    ("3x3", 260, 136, 13, 3, dict(img=L.IMG_PATCHES, patch=4)),
    # halo-tile kernel corner cases: 4 x 48 accumulator columns per tile (only two TMEM buffers -> two MMA warps / epilogue
    # groups active), and 32-channel tensors with a residual (256-bit epilogue loads / stores, two chunks per pixel)
    ("up", 40, 56, 64, 48, dict(act=True, border=L.BORDER_REPLICATE)),
    ("3x3", 90, 70, 32, 32, dict(raw=True, act=True, res=0, border=L.BORDER_REPLICATE)),
    ("3x3", 50, 38, 32, 32, dict(raw=True, res=1, border=L.BORDER_NONE)),
    # CTA-pair kernel: odd tile counts (the last pair's second CTA idles), K = 104 of 128 (seven of eight k-steps), two column blocks,
    # every ring depth (K = 32 / 64 / 128: four / two / one tile in flight)
    ("3x3", 136, 120, 104, 52, dict(raw=True, border=L.BORDER_NONE)),
    ("3x3", 72, 56, 104, 104, dict(raw=True, act=True, res=1, border=L.BORDER_REPLICATE)),
    ("3x3", 40, 24, 128, 208, dict(act=True, res=0, border=L.BORDER_CONSTANT)),
    ("3x3", 150, 90, 52, 26, dict(raw=True, border=L.BORDER_NONE)),
    ("3x3", 9, 7, 26, 26, dict(raw=True, res=0, border=L.BORDER_NONE)),
    ("1x1", 150, 90, 104, 52, dict(raw=True)),
    ("1x1", 33, 47, 52, 26, dict(act=True, border=L.BORDER_REPLICATE)),
]


def _make_conv(mode, H, W, cin, cout, ex, dtype, impl, seed):
    g = torch.Generator().manual_seed(seed)
    kin, kout = c_store(cin), c_store(cout)
    src = _grid(H, W, kin, dtype, g)
    src.buf[..., cin:] = 0
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    if mode == "3x3":
        w, m, s = PK.pack_conv3x3(wt, dtype), L.CONV3X3, 1
    elif mode == "1x1":
        w, m, s = PK.pack_conv1x1(wt[:, :, :1, :1].contiguous(), dtype), L.CONV1X1, 1
    else:
        w, m, s = PK.pack_upconv(wt, dtype), L.UPCONV, 2
    n_pad = w.shape[1]
    op = ConvOp(mode=m, src=src, w=w, k=kin, bias=PK.pad_vec(0.1 * torch.randn(cout, generator=g), n_pad), impl=impl,
                name=f"{mode}_{cin}_{cout}")
    oh, ow = s * H, s * W
    op.out_h, op.out_w, op.out_c = oh, ow, kout
    if "img" in ex:
        op.img_c, op.img_layout, op.patch = cout, ex["img"], ex.get("patch", 0)
        if ex["img"] == L.IMG_MERGED:
            op.out_img = torch.zeros(1, cout, oh, ow)
        else:
            P = ex["patch"]
            op.out_img = torch.zeros((oh // P) * (ow // P), cout, P, P)
        return op
    if ex.get("raw"):
        op.out_raw = Grid(torch.full((oh + 2, ow + 2, kout), 7.0).to(dtype), oh, ow, kout)
    if ex.get("act"):
        op.out_act = Grid(torch.full((oh + 2, ow + 2, kout), 7.0).to(dtype), oh, ow, kout)
        if not ex.get("relu"):
            op.scale = PK.pad_vec(1 + 0.1 * torch.randn(cout, generator=g), n_pad)
            op.shift = PK.pad_vec(0.1 * torch.randn(cout, generator=g), n_pad)
            op.leak = 0.02
        op.border = ex.get("border", L.BORDER_NONE)
    if "res" in ex:
        sh = ex["res"]
        rh, rw = (oh + (1 << sh) - 1) >> sh, (ow + (1 << sh) - 1) >> sh
        r = _grid(rh, rw, kout, dtype, g)
        op.res_kind, op.res, op.res_shift, op.res_c, op.res_h, op.res_w = L.RES_GRID, r.buf, sh, kout, rh, rw
    return op


def _conv_to_dev(op: ConvOp) -> ConvOp:
    import copy
    d = copy.copy(op)
    mv = lambda t: None if t is None else t.clone().cuda()
    mg = lambda g: None if g is None else Grid(g.buf.clone().cuda(), g.h, g.w, g.c)
    d.src, d.w, d.bias, d.scale, d.shift = mg(op.src), mv(op.w), mv(op.bias), mv(op.scale), mv(op.shift)
    d.out_raw, d.out_act, d.out_img, d.res = mg(op.out_raw), mg(op.out_act), mv(op.out_img), mv(op.res)
    d.mod_x, d.mod_mean, d.mod_rstd = mg(op.mod_x), mv(op.mod_mean), mv(op.mod_rstd)
    return d


def _check_conv(opc: ConvOp, opg: ConvOp, dtype):
    if opc.out_img is not None:
        err = (opg.out_img.cpu() - opc.out_img).abs().max().item()
        assert err <= (1e-5 if dtype == torch.float32 else REL[dtype] * 4), f"{opc.name}: image max-abs {err:.3e}"
        return
    bordered = opc.border != L.BORDER_NONE
    if opc.out_raw is not None:
        _cmp(opc.name + ".raw", opg.out_raw.interior, opc.out_raw.interior, dtype)
    if opc.out_act is not None:
        a, b = (opg.out_act.buf, opc.out_act.buf) if bordered else (opg.out_act.interior, opc.out_act.interior)
        _cmp(opc.name + ".act", a, b, dtype)


@pytest.mark.parametrize("precision,impl", [("fp32", L.IMPL_DIRECT), ("fp32", L.IMPL_SPLIT), ("fp16", L.IMPL_DIRECT), ("fp16", L.IMPL_UMMA),
                                            ("bf16", L.IMPL_UMMA), ("fp16", L.IMPL_TILE), ("bf16", L.IMPL_TILE),
                                            ("fp16", L.IMPL_PAIR), ("bf16", L.IMPL_PAIR)],
                         ids=["fp32-direct", "fp32-split", "fp16-direct", "fp16-umma", "bf16-umma", "fp16-tile", "bf16-tile", "fp16-pair", "bf16-pair"])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: f"{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}")
def test_conv_matches_emulator(be, case, precision, impl):
    mode, H, W, cin, cout, ex = case
    dtype = DT[precision]
    if impl == L.IMPL_TILE and (cin > 64 or cout > 64):
        pytest.skip("halo-tile kernel serves k_pad <= 64, n_pad <= 64")
    if impl == L.IMPL_PAIR and (mode == "up" or "img" in ex or k_pad_of(c_store(cin)) > 128 or cout > 256):
        pytest.skip("CTA-pair kernel serves 3x3 / 1x1 grid-to-grid convs with k_pad <= 128")
    opc = _make_conv(mode, H, W, cin, cout, ex, dtype, impl, seed=H * 1000 + W * 10 + cin + cout)
    opg = _run_both(be, opc, _conv_to_dev)
    _check_conv(opc, opg, dtype)


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("H,W,c2,c3,cout,ex", [
    (50, 38, 26, 52, 26, dict(raw=True, border=L.BORDER_NONE)),                       # block 5 of the 34 Generator: conv2 26 -> 26 + shortcut 52 -> 26
    (33, 70, 52, 104, 52, dict(raw=True, border=L.BORDER_NONE)),                      # block 4: two ring slots
    (136, 120, 26, 52, 26, dict(act=True, border=L.BORDER_REPLICATE)),                # last block: activated, framed output
    (9, 7, 16, 32, 16, dict(raw=True, act=True, res=0, border=L.BORDER_CONSTANT)),    # with a residual on top
])
def test_conv_with_folded_shortcut_matches_emulator(be, precision, H, W, c2, c3, cout, ex):
    """itg_conv_desc.in2: conv3x3(in) + conv1x1(in2) in one accumulator (the block's 1x1 shortcut folded into conv2, CTA-pair kernel)."""
    dtype = DT[precision]
    opc = _make_conv("3x3", H, W, c2, cout, ex, dtype, L.IMPL_AUTO, seed=H * 100 + W + c3)
    g = torch.Generator().manual_seed(c3 * 31 + H)
    k3 = c_store(c3)
    in2 = _grid(H, W, k3, dtype, g)
    in2.buf[..., c3:] = 0
    opc.in2, opc.k2 = in2, k3
    opc.w2 = PK.pack_conv1x1(torch.randn(cout, c3, 1, 1, generator=g) / math.sqrt(c3), dtype)

    def to_dev(op):
        d = _conv_to_dev(op)
        d.in2, d.w2 = Grid(op.in2.buf.clone().cuda(), H, W, k3), op.w2.clone().cuda()
        return d
    opg = _run_both(be, opc, to_dev)
    _check_conv(opc, opg, dtype)


@pytest.mark.parametrize("precision,impl", [("fp32", L.IMPL_DIRECT), ("fp32", L.IMPL_SPLIT), ("fp16", L.IMPL_UMMA), ("fp16", L.IMPL_TILE),
                                            ("fp16", L.IMPL_PAIR)], ids=["direct", "split", "umma", "tile", "pair"])
@pytest.mark.parametrize("mode,H,W,cin,cout", [("3x3", 33, 47, 52, 26), ("1x1", 17, 9, 64, 64), ("3x3", 130, 21, 26, 13)])
def test_conv_writes_stay_inside_their_tensors(be, precision, impl, mode, H, W, cin, cout):
    """Partial tiles, odd tile counts, padded GEMM columns: nothing is written outside the output tensors (guard zones on both sides stay
    intact; compute-sanitizer is not available on the GPU pool)."""
    dtype = DT[precision]
    opc = _make_conv(mode, H, W, cin, cout, dict(raw=True, act=True, res=0, border=L.BORDER_REPLICATE), dtype, impl, seed=H + W + cin)
    op = _conv_to_dev(opc)
    guard, canary = 4096, 123.0
    outs = {}
    for name in ("out_raw", "out_act"):
        g = getattr(op, name)
        n = g.buf.numel()
        big = torch.full((n + 2 * guard,), canary, dtype=dtype, device="cuda")
        setattr(op, name, Grid(big[guard:guard + n].view_as(g.buf), g.h, g.w, g.c))
        outs[name] = big
    be.conv(op)
    torch.cuda.synchronize()
    for name, big in outs.items():
        assert bool((big[:guard] == canary).all()) and bool((big[-guard:] == canary).all()), f"{name}: write outside the tensor"
    EmulatorBackend().conv(opc)
    _check_conv(opc, op, dtype)


@pytest.mark.parametrize("cin,cout", [(104, 52), (52, 26), (26, 26), (208, 104), (13, 13)])
def test_auto_kernel_choice_does_not_depend_on_the_grid_size(be, cin, cout):
    """A row band and the whole texture must run the same kernel for the same layer: the same pixels computed as part of a large grid and
    of a small window around them are bit-identical (two kernels would accumulate in two orders)."""
    g = torch.Generator().manual_seed(cin * 7 + cout)
    kin, kout = c_store(cin), c_store(cout)
    H, W, wy, wx, wh, ww = 384, 400, 160, 200, 24, 40            # 1 200 tiles of 16 x 8 against 15
    big = _grid(H, W, kin, torch.float16, g)
    big.buf[..., cin:] = 0
    w = PK.pack_conv3x3(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin), torch.float16).cuda()
    bias = PK.pad_vec(0.1 * torch.randn(cout, generator=g), w.shape[1]).cuda()

    def run(src_buf, h, wd):
        src = Grid(src_buf.contiguous().clone().cuda(), h, wd, kin)
        op = ConvOp(mode=L.CONV3X3, src=src, w=w, k=kin, bias=bias, impl=L.IMPL_AUTO, name="size")
        op.out_h, op.out_w, op.out_c = h, wd, kout
        op.out_raw = Grid(torch.zeros(h + 2, wd + 2, kout, dtype=torch.float16, device="cuda"), h, wd, kout)
        be.conv(op)
        torch.cuda.synchronize()
        return op.out_raw.interior.cpu()

    full = run(big.buf, H, W)
    win = run(big.buf[wy:wy + wh + 2, wx:wx + ww + 2], wh, ww)       # framed window: its frame = the neighbours in the large grid
    assert torch.equal(win, full[wy:wy + wh, wx:wx + ww])


@pytest.mark.parametrize("precision,impl", [("fp32", L.IMPL_DIRECT), ("fp32", L.IMPL_SPLIT), ("fp16", L.IMPL_AUTO), ("bf16", L.IMPL_AUTO)])
@pytest.mark.parametrize("C,shift,linear", [(52, 1, False), (26, 0, False), (104, 1, True), (8, 0, False)])
def test_ssm_embed_conv_matches_emulator(be, precision, impl, C, shift, linear):
    """The SSM pair: valid conv on a window of the hidden map + modulation epilogue (layers.py:228-234)."""
    dtype = DT[precision]
    g = torch.Generator().manual_seed(C * 7 + shift)
    H, W = 22, 36
    cs = c_store(C)
    m1 = _grid(H + 2, W + 2, 128, dtype, g)                       # interior (H+2)x(W+2): the valid conv's input
    m1.buf.clamp_(min=0)
    wt = torch.randn(2 * C, 128, 3, 3, generator=g) / math.sqrt(9 * 128)
    bias = 0.1 * torch.randn(2 * C, generator=g)
    w, b = PK.pack_ssm_embed(wt, bias, dtype)
    xh, xw = (H + shift) >> shift, (W + shift) >> shift
    x = _grid(xh, xw, cs, dtype, g)
    x.buf[..., C:] = 0
    out = Grid(torch.full((H + 2, W + 2, cs), 7.0).to(dtype), H, W, cs)
    op = ConvOp(mode=L.CONV3X3, src=m1, w=w, k=128, bias=b, impl=impl, name=f"ssm{C}")
    op.in_h, op.in_w, op.in_pitch, op.in_elem_off = H, W, W + 4, ((W + 4) + 1) * 128
    op.out_h, op.out_w, op.out_c = H, W, cs
    op.mod_x, op.mod_shift = x, shift
    op.mod_mean = PK.pad_vec(0.1 * torch.randn(C, generator=g), cs)
    op.mod_rstd = PK.pad_vec(1 + 0.2 * torch.rand(C, generator=g), cs)
    op.out_act, op.leak, op.act_linear = out, 0.02, linear
    op.border = L.BORDER_NONE if linear else L.BORDER_REPLICATE
    opg = _run_both(be, op, _conv_to_dev)
    a, r = (opg.out_act.interior, op.out_act.interior) if linear else (opg.out_act.buf, op.out_act.buf)
    _cmp(op.name, a, r, dtype)


SSM_CASES = [
    # (C, H, W, x_shift, linear, border): N = 2 * c_store(C) GEMM columns -> ceil(N / 64) resident column blocks
    (26, 22, 36, 0, False, L.BORDER_REPLICATE),        # one block, partial tiles on both edges
    (52, 22, 36, 1, False, L.BORDER_REPLICATE),        # two blocks (64 + 48 columns)
    (104, 40, 24, 1, True, L.BORDER_NONE),             # four blocks, the shortcut's bn3 (no activation, frame left to the caller)
    (8, 9, 13, 0, False, L.BORDER_CONSTANT),           # 16 columns, a single partial tile row
    (416, 12, 20, 0, False, L.BORDER_REPLICATE),       # 13 blocks, fewer tiles than SMs per block
    (26, 320, 328, 0, False, L.BORDER_REPLICATE),      # 820 tiles: 5-6 per CTA, every mbarrier wraps its phase several times
    (52, 200, 264, 1, False, L.BORDER_CONSTANT),       # two blocks x 6 tiles per CTA
]


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("case", SSM_CASES, ids=lambda c: f"C{c[0]}_{c[1]}x{c[2]}_s{c[3]}")
def test_fused_ssm_matches_emulator(be, precision, case):
    """itg_ssm_fwd (mlp_shared -> ReLU -> embed -> modulation in one launch, hidden map on chip; layers.py:228-234) against the
    CPU emulation of the same descriptor on identical operands."""
    C, H, W, shift, linear, border = case
    dtype = DT[precision]
    g = torch.Generator().manual_seed(C * 13 + H + shift)
    cs = c_store(C)
    w_mlp = PK.pack_ssm_mlp(torch.randn(128, 1, 3, 3, generator=g) / 3.0, 0.2 * torch.randn(128, generator=g), dtype)
    w_emb, b_emb = PK.pack_ssm_embed(torch.randn(2 * C, 128, 3, 3, generator=g) / math.sqrt(9 * 128), 0.1 * torch.randn(2 * C, generator=g), dtype)
    xh, xw = (H + shift) >> shift, (W + shift) >> shift
    x = _grid(xh, xw, cs, dtype, g)
    x.buf[..., C:] = 0
    mk = lambda dev: SsmOp(
        map=torch.randn(H + 4, W + 4, generator=torch.Generator().manual_seed(H * W)).to(dev), w_mlp=w_mlp.to(dev), w_embed=w_emb.to(dev),
        b_embed=b_emb.to(dev), x=Grid(x.buf.clone().to(dev), xh, xw, cs), x_shift=shift,
        mean=PK.pad_vec(0.1 * torch.randn(C, generator=torch.Generator().manual_seed(1)), cs).to(dev),
        rstd=PK.pad_vec(1 + 0.2 * torch.rand(C, generator=torch.Generator().manual_seed(2)), cs).to(dev),
        out=Grid(torch.full((H + 2, W + 2, cs), 7.0).to(dtype).to(dev), H, W, cs), leak=0.02, linear=linear, border=border, name=f"ssm{C}")
    opc, opg = mk("cpu"), mk("cuda")
    EmulatorBackend().ssm(opc)
    be.ssm(opg)
    torch.cuda.synchronize()
    a, r = (opg.out.interior, opc.out.interior) if border == L.BORDER_NONE else (opg.out.buf, opc.out.buf)
    # the hidden map is rounded to the operand type on both sides; a rounding flip there (bias carried as hi + lo) moves one of the
    # 1152 products by one ulp: allow 2x the single-rounding budget
    got, ref = a.float().cpu(), r.float()
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item()
    assert err <= 2 * REL[dtype] * scale + 1e-6, f"ssm C={C}: max-abs {err:.3e} at output scale {scale:.3e}"
    with pytest.raises(L.ItgError):
        bad = mk("cuda")
        bad.map = bad.map[:, :-1].contiguous()
        be.ssm(bad)


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("C,patch,th,tw", [(104, 16, 2, 3), (16, 16, 3, 2), (128, 16, 1, 2), (32, 8, 2, 2)])
def test_attention_matches_emulator(be, precision, C, patch, th, tw):
    dtype = DT[precision]
    g = torch.Generator().manual_seed(C + patch)
    xc = c_store(C)
    x = _grid(th * patch, tw * patch, xc, dtype, g)
    x.buf[..., C:] = 0
    rnd = lambda *s: torch.randn(*s, generator=g)
    kw = dict(w_theta=rnd(C // 8, C) / math.sqrt(C), b_theta=0.1 * rnd(C // 8), w_phi=rnd(C // 8, C) / math.sqrt(C),
              b_phi=0.1 * rnd(C // 8), w_g=rnd(C // 2, C) / math.sqrt(C), b_g=0.1 * rnd(C // 2),
              w_o=rnd(C, C // 2) / math.sqrt(C // 2), b_o=0.1 * rnd(C), gamma=torch.tensor([0.7]),
              scale=PK.pad_vec(1 + 0.1 * rnd(C), xc), shift=PK.pad_vec(0.1 * rnd(C), xc))
    mk = lambda dev: AttentionOp(
        x=Grid(x.buf.clone().to(dev), x.h, x.w, xc), th=th, tw=tw, patch=patch, C=C,
        out_raw=Grid(torch.full_like(x.buf, 7.0).to(dev), x.h, x.w, xc), out_act=Grid(torch.full_like(x.buf, 7.0).to(dev), x.h, x.w, xc),
        leak=0.02, border=L.BORDER_REPLICATE, **{k: v.to(dev) for k, v in kw.items()})
    opc, opg = mk("cpu"), mk("cuda")
    EmulatorBackend().attention(opc)
    be.attention(opg)
    torch.cuda.synchronize()
    _cmp("att.raw", opg.out_raw.interior, opc.out_raw.interior, dtype)
    _cmp("att.act", opg.out_act.buf, opc.out_act.buf, dtype)


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_data_movement_is_bit_exact(be, precision):
    dtype = DT[precision]
    g = torch.Generator().manual_seed(3)
    emu = EmulatorBackend()
    # pack_nchw (z grid, utils.py:228)
    src = torch.randn(13, 10, 18, generator=g)
    dc, dg = torch.full((10, 18, 16), 7.0).to(dtype), torch.full((10, 18, 16), 7.0).to(dtype).cuda()
    emu.pack_nchw(src, dc)
    be.pack_nchw(src.cuda(), dg)
    assert torch.equal(dg.cpu(), dc)
    # pack_map_taps (SSM noise map, utils.py:246)
    m = torch.randn(14, 23, generator=g)
    tc = Grid(torch.zeros(14, 23, 16).to(dtype), 12, 21, 16)
    tg = Grid(torch.zeros(14, 23, 16).to(dtype).cuda(), 12, 21, 16)
    emu.pack_map_taps(m, tc)
    be.pack_map_taps(m.cuda(), tg)
    assert torch.equal(tg.interior.cpu(), tc.interior)
    # copy_rect / fill_frame (halo moves, F.pad of layers.py:82)
    a = _grid(9, 11, 24, dtype, g)
    for border in (L.BORDER_REPLICATE, L.BORDER_CONSTANT):
        for sides in (15, 1, 2, 4, 8, 5, 10):
            c, d = _clone_grid(a, "cpu"), _clone_grid(a, "cuda")
            emu.fill_frame(c, border, sides)
            be.fill_frame(d, border, sides)
            assert torch.equal(d.buf.cpu(), c.buf), (border, sides)
    dst_c, dst_g = torch.zeros(5, 30, 24).to(dtype), torch.zeros(5, 30, 24).to(dtype).cuda()
    emu.copy_rect(a.buf, 2, 3, dst_c, 1, 4, 3, 7)
    be.copy_rect(a.buf.cuda(), 2, 3, dst_g, 1, 4, 3, 7)
    assert torch.equal(dst_g.cpu(), dst_c)


def test_bad_arguments_raise(be):
    """Errors come back as ItgError with the library's message, not as crashes."""
    g = torch.Generator().manual_seed(1)
    op = _conv_to_dev(_make_conv("3x3", 8, 8, 16, 16, dict(raw=True), torch.float16, L.IMPL_UMMA, 1))
    op.out_h = 9
    with pytest.raises(L.ItgError, match="does not match"):
        be.conv(op)
    with pytest.raises(L.ItgError, match="CUDA tensors"):
        L.ptr(torch.zeros(4))


@pytest.mark.parametrize("mode_name,border", [("replicate", L.BORDER_REPLICATE), ("constant", L.BORDER_CONSTANT)])
def test_halo_gather_indexing_is_bit_exact_vs_reference_local_padder(be, mode_name, border):
    """The conv kernels' neighbourhood gather == models.layers.LocalPadder windows, bit for bit.

    tests/golden/localpad.npz holds integer-coded patch batches and the (B, C, r+2, r+2) windows the UNMODIFIED
    reference LocalPadder produced for them.  A 3x3 conv whose weight for output column t*C+c is the delta at tap t,
    channel c copies the gathered neighbourhood to its output: out[y, x, t*C + c] = window[c, y%r + t//3, x%r + t%3].
    fp32 CUDA-core and split-precision kernels on the golden integers; tcgen05 halo-tile, streaming and CTA-pair kernels (fp16 holds
    integers < 2048 exactly) on the same values reduced mod 2039, all compared exactly."""
    import os
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "localpad.npz"))
    x = torch.from_numpy(d[f"{mode_name}_0_in"])            # first sub-image: pure outer padding ('1st_row_1st_col')
    win = torch.from_numpy(d[f"{mode_name}_0_out"])          # (9, C, r+2, r+2)
    B, C, r, _ = x.shape
    nph = npw = 3
    H, W = nph * r, npw * r
    merged = x.reshape(nph, npw, C, r, r).permute(2, 0, 3, 1, 4).reshape(C, H, W)
    # delta weights: (9*C, C, 3, 3)
    wt = torch.zeros(9 * C, C, 3, 3)
    for t in range(9):
        for c in range(C):
            wt[t * C + c, c, t // 3, t % 3] = 1.0
    ref = torch.empty(H, W, 9 * C)
    for p in range(B):
        py, px = p // npw, p % npw
        for t in range(9):
            ref[py * r:(py + 1) * r, px * r:(px + 1) * r, t * C:(t + 1) * C] = \
                win[p, :, t // 3:t // 3 + r, t % 3:t % 3 + r].permute(1, 2, 0)
    for precision, impl, modulo in (("fp32", L.IMPL_DIRECT, None), ("fp32", L.IMPL_SPLIT, None), ("fp16", L.IMPL_TILE, 2039),
                                    ("fp16", L.IMPL_UMMA, 2039), ("fp16", L.IMPL_PAIR, 2039), ("bf16", L.IMPL_TILE, 251)):
        dtype = DT[precision]
        src_v = merged if modulo is None else torch.remainder(merged, modulo)
        ref_v = ref if modulo is None else torch.remainder(ref, modulo)
        if modulo is not None and border == L.BORDER_CONSTANT:
            ref_v = torch.where(ref == 0, torch.zeros_like(ref_v), ref_v)        # padding zeros stay zeros
        kin = c_store(C)
        buf = torch.full((H + 2, W + 2, kin), -7.0)
        buf[1:-1, 1:-1, :C] = src_v.permute(1, 2, 0)
        buf[1:-1, 1:-1, C:] = 0
        g = Grid(buf.to(dtype).cuda(), H, W, kin)
        be.fill_frame(g, border)                                                   # F.pad of layers.py:82
        w = PK.pack_conv3x3(wt, dtype).cuda()
        out = Grid(torch.zeros((H + 2, W + 2, c_store(9 * C)), dtype=dtype, device="cuda"), H, W, c_store(9 * C))
        op = ConvOp(mode=L.CONV3X3, src=g, w=w, k=kin, bias=None, impl=impl, name="gather")
        op.out_h, op.out_w, op.out_c, op.out_raw = H, W, c_store(9 * C), out
        be.conv(op)
        torch.cuda.synchronize()
        got = out.interior[..., :9 * C].float().cpu()
        assert torch.equal(got, ref_v), (precision, impl, (got - ref_v).abs().max().item())
