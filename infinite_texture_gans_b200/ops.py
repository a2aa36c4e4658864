"""Grid tensors and the operator layer over libitg_b200.so.

A *grid tensor* is the merged, channels-last form of the reference's patch batch (include/itg.h, "Data
layout"): interior H x W pixels, C storage channels, stored with a 1-pixel frame as a contiguous
(H+2, W+2, C) buffer.  `ConvOp` / `AttentionOp` describe one fused launch with tensors instead of raw
pointers; `CudaBackend` turns them into C-ABI calls.  The engine (engine.py) only talks to a backend
object, which is what lets the CPU test-suite substitute an emulator for the *host logic* tests --
the package itself ships no other backend.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib as L


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def c_store(c: int) -> int:
    """Storage channels of a grid tensor holding c real channels."""
    return round_up(c, 8)


def n_pad_of(n: int) -> int:
    """GEMM columns for n real output columns: multiple of 16 that splits evenly into <=256-wide blocks."""
    n16 = round_up(n, 16)
    nb = (n16 + 255) // 256
    return round_up(n16, 16 * nb)


def pair_fold_eligible(k: int, k2: int, n: int) -> bool:
    """Can a 3x3 conv with k stored input channels take a folded 1x1 shortcut over k2 stored channels (itg_conv_desc.in2)?  Mirror of
    pair_eligible() in csrc/itg_abi.cu: both inputs of a 16x8 tile within half the activation ring (45 planes) and within the loaders'
    16 x 192 sixteen-byte chunks; <= 128 channels each; <= 256 GEMM columns."""
    if k > 128 or k2 > 128 or n_pad_of(n) > 256:
        return False
    planes = 2 * ((k // 8 + 1) // 2) + 2 * ((k2 // 8 + 1) // 2)
    chunks = 180 * (k // 8) + 128 * (k2 // 8)
    return 2 * planes <= 45 and chunks <= 16 * 192


def k_pad_of(k: int) -> int:
    """Padded contraction length per tap: one 32/64-byte swizzle row, or whole 128-byte rows."""
    return 16 if k <= 16 else (32 if k <= 32 else round_up(k, 64))


@dataclass
class Grid:
    """Framed channels-last tensor: buf is (h+2, w+2, c) contiguous (possibly a view into a pooled block)."""
    buf: torch.Tensor
    h: int
    w: int
    c: int

    @staticmethod
    def empty(h: int, w: int, c: int, dtype, device, zero: bool = False) -> "Grid":
        f = torch.zeros if zero else torch.empty
        return Grid(f((h + 2, w + 2, c), dtype=dtype, device=device), h, w, c)

    @property
    def interior(self) -> torch.Tensor:
        return self.buf[1:-1, 1:-1]

    def to_nchw(self, channels: Optional[int] = None) -> torch.Tensor:
        """Interior as a (1, C, H, W) fp32 tensor (tests / debugging)."""
        t = self.interior[..., : (channels or self.c)]
        return t.permute(2, 0, 1).unsqueeze(0).float().contiguous()


@dataclass
class ConvOp:
    """One itg_conv_fwd launch; tensor-valued mirror of itg_conv_desc."""
    mode: int
    src: Grid
    w: torch.Tensor                    # [taps, n_pad, k_pad]
    k: int
    bias: Optional[torch.Tensor] = None
    in_c_off: int = 0
    # window into a wider buffer (SSM hidden map): interior size / pitch override
    in_h: Optional[int] = None
    in_w: Optional[int] = None
    in_pitch: int = 0
    in_elem_off: int = 0               # element offset of the window's frame origin inside src.buf
    out_h: int = 0
    out_w: int = 0
    out_c: int = 0
    border: int = L.BORDER_NONE
    res_kind: int = L.RES_NONE
    res: Optional[torch.Tensor] = None  # Grid.buf or fp32 (h, w, c)
    res_shift: int = 0
    res_c: int = 0
    res_h: int = 0
    res_w: int = 0
    mod_x: Optional[Grid] = None
    mod_shift: int = 0
    mod_mean: Optional[torch.Tensor] = None
    mod_rstd: Optional[torch.Tensor] = None
    out_raw: Optional[Grid] = None
    out_act: Optional[Grid] = None
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    leak: float = 0.0
    act_linear: bool = False
    out_f32: Optional[torch.Tensor] = None
    out_img: Optional[torch.Tensor] = None
    img_c: int = 0
    img_layout: int = L.IMG_MERGED
    patch: int = 0
    impl: int = L.IMPL_AUTO
    name: str = ""
    # optional second input: the block's 1x1 shortcut folded into this 3x3 conv (itg_conv_desc.in2; CTA-pair kernel only)
    in2: Optional[Grid] = None
    w2: Optional[torch.Tensor] = None   # [1, n_pad, k2_pad]
    k2: int = 0

    @property
    def m_h(self) -> int:
        return self.in_h if self.in_h is not None else self.src.h

    @property
    def m_w(self) -> int:
        return self.in_w if self.in_w is not None else self.src.w


@dataclass
class AttentionOp:
    x: Grid
    th: int
    tw: int
    patch: int
    C: int
    w_theta: torch.Tensor
    b_theta: torch.Tensor
    w_phi: torch.Tensor
    b_phi: torch.Tensor
    w_g: torch.Tensor
    b_g: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    gamma: torch.Tensor
    out_raw: Optional[Grid] = None
    out_act: Optional[Grid] = None
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    leak: float = 0.0
    border: int = L.BORDER_NONE
    name: str = "attention"


@dataclass
class SsmOp:
    """One itg_ssm_fwd launch (StochasticSpatialModulation.forward, models/layers.py:228-234); mirror of itg_ssm_desc."""
    map: torch.Tensor                  # fp32 (h+4, w+4) noise map of the level
    w_mlp: torch.Tensor                # [128, 16] packing.pack_ssm_mlp
    w_embed: torch.Tensor              # [9, n_pad, 128] packing.pack_ssm_embed
    b_embed: torch.Tensor              # [n_pad] fp32
    x: Grid                            # tensor to modulate
    x_shift: int
    mean: torch.Tensor
    rstd: torch.Tensor
    out: Grid
    leak: float = 0.0
    linear: bool = False
    border: int = L.BORDER_NONE
    zero_ring: bool = False            # non-local Generator: hidden map zero outside the image
    name: str = "ssm"


class CudaBackend:
    """Executes ops through libitg_b200.so on the current CUDA stream.  The only backend of the package."""

    name = "cuda"

    def __init__(self):
        self.lib = L.load()
        self.launches = 0          # kernels launched through this backend (bench.py's gpu_launches)

    # ---- conv ----
    def _desc(self, op: ConvOp) -> L.ConvDesc:
        d = L.ConvDesc()
        src = op.src
        dt = L.DTYPE_OF[src.buf.dtype]
        d.dtype, d.mode, d.impl, d.border = dt, op.mode, op.impl, op.border
        d.in_ = L.ptr(src.buf) + op.in_elem_off * src.buf.element_size()
        d.in_h, d.in_w, d.in_pitch = op.m_h, op.m_w, op.in_pitch
        d.in_c, d.in_c_off, d.k = src.c, op.in_c_off, op.k
        if op.w.dtype != src.buf.dtype or not op.w.is_contiguous():
            raise L.ItgError(f"conv {op.name}: weights must be contiguous and of the activations' dtype")
        d.w, d.n_pad, d.k_pad = L.ptr(op.w), op.w.shape[1], op.w.shape[2]
        d.bias = L.ptr(op.bias)
        d.out_h, d.out_w, d.out_c = op.out_h, op.out_w, op.out_c
        d.res_kind, d.res_shift, d.res = op.res_kind, op.res_shift, L.ptr(op.res)
        d.res_c, d.res_h, d.res_w = op.res_c, op.res_h, op.res_w
        if op.mod_x is not None:
            d.mod_x, d.mod_c, d.mod_shift = L.ptr(op.mod_x.buf), op.mod_x.c, op.mod_shift
            d.mod_h, d.mod_w = op.mod_x.h, op.mod_x.w
            d.mod_mean, d.mod_rstd = L.ptr(op.mod_mean), L.ptr(op.mod_rstd)
        d.out_raw = L.ptr(op.out_raw.buf) if op.out_raw is not None else None
        d.out_act = L.ptr(op.out_act.buf) if op.out_act is not None else None
        d.scale, d.shift, d.leak, d.act_linear = L.ptr(op.scale), L.ptr(op.shift), float(op.leak), int(op.act_linear)
        d.out_f32, d.out_img = L.ptr(op.out_f32), L.ptr(op.out_img)
        d.img_c, d.img_layout, d.patch = op.img_c, op.img_layout, op.patch
        if op.in2 is not None:
            if op.w2 is None or op.w2.dtype != src.buf.dtype or not op.w2.is_contiguous() or op.w2.shape[1] != op.w.shape[1]:
                raise L.ItgError(f"conv {op.name}: w2 must be a contiguous [1, n_pad, k2_pad] tensor of the activations' dtype")
            if (op.in2.h, op.in2.w) != (op.m_h, op.m_w) or op.in2.buf.dtype != src.buf.dtype:
                raise L.ItgError(f"conv {op.name}: the second input must have the first one's interior size and dtype")
            d.in2, d.in2_c, d.in2_c_off, d.k2 = L.ptr(op.in2.buf), op.in2.c, 0, op.k2 or op.in2.c
            d.w2, d.k2_pad = L.ptr(op.w2), op.w2.shape[2]
        return d

    def compile_conv(self, op: ConvOp):
        """Pre-build the descriptor; the returned closure is one itg_conv_fwd launch on the current stream."""
        d, fn, name = self._desc(op), self.lib.itg_conv_fwd, op.name

        def launch():
            rc = fn(d, L.stream_ptr())
            if rc != 0:
                raise L.ItgError(f"conv {name}: libitg_b200 error {rc}: {self.lib.itg_last_error().decode()}")
            self.launches += 1
        return launch

    def conv(self, op: ConvOp) -> None:
        self.compile_conv(op)()

    # ---- SSM ----
    def compile_ssm(self, op: SsmOp):
        d = L.SsmDesc()
        out = op.out
        if op.map.dtype != torch.float32 or op.map.dim() != 2 or op.map.stride(1) != 1:
            raise L.ItgError(f"ssm {op.name}: the noise map must be a 2-D fp32 tensor with unit column stride")
        if tuple(op.map.shape) != (out.h + 4, out.w + 4):
            raise L.ItgError(f"ssm {op.name}: map {tuple(op.map.shape)} does not match the {out.h}x{out.w} output (+4)")
        if op.w_mlp.dtype != out.buf.dtype or op.w_embed.dtype != out.buf.dtype or op.x.buf.dtype != out.buf.dtype:
            raise L.ItgError(f"ssm {op.name}: weights and tensors must share one 16-bit dtype")
        if tuple(op.w_mlp.shape) != (128, 16) or op.w_embed.shape[0] != 9 or op.w_embed.shape[2] != 128 or not op.w_embed.is_contiguous():
            raise L.ItgError(f"ssm {op.name}: weight layouts must be [128,16] and [9,n_pad,128]")
        d.dtype, d.border, d.h, d.w, d.c, d.n_pad = L.DTYPE_OF[out.buf.dtype], op.border, out.h, out.w, out.c, op.w_embed.shape[1]
        d.map, d.map_pitch, d.x_shift = L.ptr(op.map), op.map.stride(0), op.x_shift
        d.w_mlp, d.w_embed, d.b_embed = L.ptr(op.w_mlp), L.ptr(op.w_embed), L.ptr(op.b_embed)
        d.x, d.x_c, d.x_h, d.x_w, d.linear = L.ptr(op.x.buf), op.x.c, op.x.h, op.x.w, int(op.linear)
        d.mean, d.rstd, d.out, d.leak = L.ptr(op.mean), L.ptr(op.rstd), L.ptr(out.buf), float(op.leak)
        d.zero_ring = int(op.zero_ring)
        fn, name = self.lib.itg_ssm_fwd, op.name

        def launch():
            rc = fn(d, L.stream_ptr())
            if rc != 0:
                raise L.ItgError(f"ssm {name}: libitg_b200 error {rc}: {self.lib.itg_last_error().decode()}")
            self.launches += 1
        return launch

    def ssm(self, op: SsmOp) -> None:
        self.compile_ssm(op)()

    # ---- attention ----
    def compile_attention(self, op: AttentionOp):
        x = op.x
        args = (L.DTYPE_OF[x.buf.dtype], L.ptr(x.buf), op.th, op.tw, op.patch, op.C, x.c,
                L.ptr(op.w_theta), L.ptr(op.b_theta), L.ptr(op.w_phi), L.ptr(op.b_phi), L.ptr(op.w_g), L.ptr(op.b_g),
                L.ptr(op.w_o), L.ptr(op.b_o), L.ptr(op.gamma),
                L.ptr(op.out_raw.buf) if op.out_raw is not None else None,
                L.ptr(op.out_act.buf) if op.out_act is not None else None,
                L.ptr(op.scale), L.ptr(op.shift), float(op.leak), op.border)

        def launch():
            L.check(self.lib.itg_attention_fwd(*args, L.stream_ptr()))
            self.launches += 1
        return launch

    def attention(self, op: AttentionOp) -> None:
        self.compile_attention(op)()

    def compile_pack_nchw(self, src: torch.Tensor, dst: torch.Tensor):
        return lambda: self.pack_nchw(src, dst)

    def compile_pack_map_taps(self, src: torch.Tensor, dst: Grid):
        return lambda: self.pack_map_taps(src, dst)

    # ---- data movement ----
    def pack_nchw(self, src: torch.Tensor, dst: torch.Tensor) -> None:
        """src fp32 (C,H,W) -> dst (H,W,dst_c) channels-last."""
        Cc, H, W = src.shape
        assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
        assert tuple(dst.shape[:2]) == (H, W)
        L.check(self.lib.itg_pack_nchw(L.DTYPE_OF[dst.dtype], L.ptr(src), Cc, H, W, L.ptr(dst), dst.shape[2], L.stream_ptr()))
        self.launches += 1

    def pack_map_taps(self, src: torch.Tensor, dst: Grid) -> None:
        """src fp32 (Hm, Wm) noise map -> dst grid with interior (Hm-2, Wm-2): channel t = map shifted by 3x3 tap t."""
        Hm, Wm = src.shape
        assert src.dtype == torch.float32 and src.is_contiguous()
        assert dst.h == Hm - 2 and dst.w == Wm - 2
        L.check(self.lib.itg_pack_map_taps(L.DTYPE_OF[dst.buf.dtype], L.ptr(src), Hm, Wm, L.ptr(dst.buf), dst.c, L.stream_ptr()))
        self.launches += 1

    def copy_rect(self, src: torch.Tensor, sy: int, sx: int, dst: torch.Tensor, dy: int, dx: int, h: int, w: int) -> None:
        """Copy h x w pixels (all channels) between (rows, pitch, c) buffers; coordinates are buffer pixels."""
        assert src.dtype == dst.dtype and src.shape[2] == dst.shape[2]
        assert 0 <= sy and sy + h <= src.shape[0] and 0 <= sx and sx + w <= src.shape[1]
        assert 0 <= dy and dy + h <= dst.shape[0] and 0 <= dx and dx + w <= dst.shape[1]
        L.check(self.lib.itg_copy_rect(L.DTYPE_OF[src.dtype], L.ptr(src), src.shape[1], sy, sx, L.ptr(dst), dst.shape[1],
                                       dy, dx, h, w, src.shape[2], L.stream_ptr()))
        self.launches += 1

    def fill_frame(self, g: Grid, border: int, sides: int = 15) -> None:
        L.check(self.lib.itg_fill_frame(L.DTYPE_OF[g.buf.dtype], L.ptr(g.buf), g.h, g.w, g.c, border, sides, L.stream_ptr()))
        self.launches += 1
