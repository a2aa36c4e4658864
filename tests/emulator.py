"""CPU emulation of libitg_b200's launch semantics (TEST INFRASTRUCTURE ONLY -- lives in tests/, never in the package).

It restates, with plain torch CPU ops in fp32, what each C-ABI entry point of include/itg.h computes on
grid tensors, so that
  * the host logic (engine.py plans, packing.py layouts, the sequential halo protocol, the row-band
    split) can be checked against the oracle on a machine without a GPU, and
  * on the GPU box every CUDA kernel can be compared launch by launch with an independent implementation
    of the same descriptor.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from infinite_texture_gans_b200 import _lib as L
from infinite_texture_gans_b200.ops import AttentionOp, ConvOp, Grid, SsmOp


def _act(v, leak):
    return torch.where(v >= 0, v, v * leak)


def _fill_frame(buf: torch.Tensor, border: int, sides: int = 15) -> None:
    """buf (h+2, w+2, c).  sides: bit0 top, bit1 bottom, bit2 left, bit3 right (a frame pixel is written
    when every side it lies on is enabled)."""
    if border == L.BORDER_NONE:
        return
    h, w = buf.shape[0] - 2, buf.shape[1] - 2
    src = buf.clone()
    for fy in range(-1, h + 1):
        for fx in range(-1, w + 1):
            top, bot, lef, rig = fy < 0, fy >= h, fx < 0, fx >= w
            if not (top or bot or lef or rig):
                continue
            if (top and not sides & 1) or (bot and not sides & 2) or (lef and not sides & 4) or (rig and not sides & 8):
                continue
            if border == L.BORDER_REPLICATE:
                cy, cx = min(max(fy, 0), h - 1), min(max(fx, 0), w - 1)
                buf[fy + 1, fx + 1] = src[cy + 1, cx + 1]
            else:
                buf[fy + 1, fx + 1] = 0


def _fill_frame_fast(buf: torch.Tensor, border: int) -> None:
    if border == L.BORDER_NONE:
        return
    if border == L.BORDER_REPLICATE:
        buf[0, 1:-1] = buf[1, 1:-1]
        buf[-1, 1:-1] = buf[-2, 1:-1]
        buf[:, 0] = buf[:, 1]
        buf[:, -1] = buf[:, -2]
    else:
        buf[0] = 0
        buf[-1] = 0
        buf[:, 0] = 0
        buf[:, -1] = 0


def _up(t: torch.Tensor, shift: int, oh: int, ow: int) -> torch.Tensor:
    if shift:
        t = t.repeat_interleave(1 << shift, 0).repeat_interleave(1 << shift, 1)
    return t[:oh, :ow]


class EmulatorBackend:
    name = "emulator"

    def __init__(self):
        self.launches = 0

    # ---- conv (itg_conv_fwd) ----
    def conv(self, op: ConvOp) -> None:
        self.launches += 1
        src = op.src.buf
        c = op.src.c
        H, W = op.m_h, op.m_w
        pitch = op.in_pitch or (W + 2)
        flat = src.reshape(-1)
        win = torch.as_strided(flat, (H + 2, W + 2, c), (pitch * c, c, 1), flat.storage_offset() + op.in_elem_off)
        x = win[..., op.in_c_off:op.in_c_off + op.k].float()
        wt = op.w.float()[:, :, :op.k]                       # [taps, n_pad, k]
        n_pad = wt.shape[1]

        def tap(dy, dx, t):
            return x[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] @ wt[t].t()

        if op.mode == L.CONV3X3:
            acc = sum(tap(t // 3 - 1, t % 3 - 1, t) for t in range(9))
            if op.in2 is not None:                             # folded 1x1 shortcut: same accumulator
                k2 = op.k2 or op.in2.c
                acc = acc + op.in2.interior[..., :k2].float() @ op.w2.float()[0, :, :k2].t()
        elif op.mode == L.CONV1X1:
            acc = tap(0, 0, 0)
        else:
            acc = torch.zeros(2 * H, 2 * W, n_pad)
            for a in (0, 1):
                for b in (0, 1):
                    ph = a * 2 + b
                    acc[a::2, b::2] = sum(tap(a - 1 + i, b - 1 + j, ph * 4 + i * 2 + j) for i in (0, 1) for j in (0, 1))
        oh, ow = op.out_h, op.out_w
        assert tuple(acc.shape[:2]) == (oh, ow)
        v = acc + (op.bias.float() if op.bias is not None else 0)

        if op.out_img is not None:
            img = torch.tanh(v[..., :op.img_c]).permute(2, 0, 1)            # (C, H, W)
            if op.img_layout == L.IMG_PATCHES:
                P = op.patch
                img = img.reshape(op.img_c, oh // P, P, ow // P, P).permute(1, 3, 0, 2, 4).reshape(-1, op.img_c, P, P)
            op.out_img.copy_(img.reshape(op.out_img.shape))
            return

        oc = op.out_c
        if op.mod_x is not None:
            gamma, beta = v[..., 0:2 * oc:2], v[..., 1:2 * oc:2]
            xm = _up(op.mod_x.interior[..., :oc].float(), op.mod_shift, oh, ow)
            y = (1 + gamma) * ((xm - op.mod_mean.float()[:oc]) * op.mod_rstd.float()[:oc]) + beta
            if not op.act_linear:
                y = _act(y, op.leak)
            op.out_act.interior.copy_(y.to(op.out_act.buf.dtype))
            _fill_frame_fast(op.out_act.buf, op.border)
            return

        v = v[..., :oc]
        if op.res_kind == L.RES_GRID:
            res = op.res.reshape(op.res_h + 2, op.res_w + 2, op.res_c)[1:-1, 1:-1, :oc].float()
            v = v + _up(res, op.res_shift, oh, ow)
        elif op.res_kind == L.RES_F32:
            v = v + _up(op.res.reshape(op.res_h, op.res_w, op.res_c)[..., :oc].float(), op.res_shift, oh, ow)
        if op.out_raw is not None:
            op.out_raw.interior.copy_(v.to(op.out_raw.buf.dtype))
        if op.out_f32 is not None:
            op.out_f32.copy_(v)
        if op.out_act is not None:
            s = op.scale.float()[:oc] if op.scale is not None else 1.0
            t = op.shift.float()[:oc] if op.shift is not None else 0.0
            y = s * v + t
            if not op.act_linear:
                y = _act(y, op.leak)
            op.out_act.interior.copy_(y.to(op.out_act.buf.dtype))
            _fill_frame_fast(op.out_act.buf, op.border)

    def compile_conv(self, op: ConvOp):
        return lambda: self.conv(op)

    # ---- SSM (itg_ssm_fwd) ----
    def ssm(self, op: SsmOp) -> None:
        """m1 = relu(conv3x3_valid(map rounded to the operand type) + bias), rounded to the operand type; embed conv + modulation in fp32."""
        self.launches += 1
        dt = op.out.buf.dtype
        H, W, C = op.out.h, op.out.w, op.out.c
        m = op.map.to(dt).float()                                           # (H+4, W+4)
        w1 = op.w_mlp.float()                                               # [128, 16]
        taps = torch.stack([m[t // 3:t // 3 + H + 2, t % 3:t % 3 + W + 2] for t in range(9)], -1)      # (H+2, W+2, 9)
        m1 = torch.relu(taps @ w1[:, :9].t() + (w1[:, 9] + w1[:, 10])).to(dt).float()                    # (H+2, W+2, 128)
        if op.zero_ring:                                                    # zero padding of embed's input (non-local Generator)
            m1[0], m1[-1], m1[:, 0], m1[:, -1] = 0, 0, 0, 0
        we = op.w_embed.float()                                             # [9, n_pad, 128]
        v = sum(m1[t // 3:t // 3 + H, t % 3:t % 3 + W] @ we[t].t() for t in range(9)) + op.b_embed.float()
        gamma, beta = v[..., 0:2 * C:2], v[..., 1:2 * C:2]
        xm = _up(op.x.interior[..., :C].float(), op.x_shift, H, W)
        y = (1 + gamma) * ((xm - op.mean.float()[:C]) * op.rstd.float()[:C]) + beta
        if not op.linear:
            y = _act(y, op.leak)
        op.out.interior.copy_(y.to(dt))
        _fill_frame_fast(op.out.buf, op.border)

    def compile_ssm(self, op: SsmOp):
        return lambda: self.ssm(op)

    # ---- attention (itg_attention_fwd) ----
    def attention(self, op: AttentionOp) -> None:
        self.launches += 1
        C, P = op.C, op.patch
        x = op.x.interior[..., :C].float()                                    # (H, W, C)
        H, W = x.shape[:2]
        th, tw = H // P, W // P
        xp = x.reshape(th, P, tw, P, C).permute(0, 2, 4, 1, 3).reshape(th * tw, C, P, P)
        conv = lambda t, w, b: F.conv2d(t, w.float().reshape(w.shape[0], -1, 1, 1), b.float())
        theta = conv(xp, op.w_theta, op.b_theta)
        phi = F.max_pool2d(conv(xp, op.w_phi, op.b_phi), 2)
        g = F.max_pool2d(conv(xp, op.w_g, op.b_g), 2)
        B = xp.shape[0]
        beta = F.softmax(torch.bmm(theta.reshape(B, C // 8, -1).transpose(1, 2), phi.reshape(B, C // 8, -1)), -1)
        o = torch.bmm(g.reshape(B, C // 2, -1), beta.transpose(1, 2)).reshape(B, C // 2, P, P)
        o = conv(o, op.w_o, op.b_o)
        out = op.gamma.float().reshape(()) * o + xp
        out = out.reshape(th, tw, C, P, P).permute(0, 3, 1, 4, 2).reshape(H, W, C)
        xc = op.x.c
        full = torch.zeros(H, W, xc)
        full[..., :C] = out
        if op.out_raw is not None:
            op.out_raw.interior.copy_(full.to(op.out_raw.buf.dtype))
        if op.out_act is not None:
            s = op.scale.float()[:xc] if op.scale is not None else 1.0
            t = op.shift.float()[:xc] if op.shift is not None else 0.0
            op.out_act.interior.copy_(_act(s * full + t, op.leak).to(op.out_act.buf.dtype))
            _fill_frame_fast(op.out_act.buf, op.border)

    def compile_attention(self, op: AttentionOp):
        return lambda: self.attention(op)

    # ---- data movement ----
    def pack_nchw(self, src: torch.Tensor, dst: torch.Tensor) -> None:
        self.launches += 1
        C = src.shape[0]
        dst.zero_()
        dst[..., :C] = src.permute(1, 2, 0).to(dst.dtype)

    def compile_pack_nchw(self, src, dst):
        return lambda: self.pack_nchw(src, dst)

    def pack_map_taps(self, src: torch.Tensor, dst: Grid) -> None:
        self.launches += 1
        h, w = dst.h, dst.w
        it = dst.interior
        it.zero_()
        for t in range(9):
            it[..., t] = src[t // 3:t // 3 + h, t % 3:t % 3 + w].to(it.dtype)

    def compile_pack_map_taps(self, src, dst):
        return lambda: self.pack_map_taps(src, dst)

    def copy_rect(self, src, sy, sx, dst, dy, dx, h, w) -> None:
        self.launches += 1
        dst[dy:dy + h, dx:dx + w] = src[sy:sy + h, sx:sx + w].clone()

    def fill_frame(self, g: Grid, border: int, sides: int = 15) -> None:
        self.launches += 1
        _fill_frame(g.buf, border, sides)
