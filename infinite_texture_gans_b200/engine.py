"""Launch plans for the Generator forward on a merged patch grid.

The reference evaluates `ResidualPatchGenerator.forward` (models/generators.py:86-124) on a batch of
patches and lets `LocalPadder` (models/layers.py:38-173) rebuild every conv's neighbourhood by merging,
padding and re-cropping.  Here the whole th x tw patch grid lives on the device as *grid tensors*
(include/itg.h) and the forward is a fixed list of fused launches:

  BN mode, block k >= 2          (ResBlockGenerator.forward, models/layers.py:301-322)
    s  = conv1x1(h)                                   shortcut at LOW resolution (commutes with nearest-up)
    a2 = act(bn2(upconv(a)))                          conv1 with the 2x up-sampling folded into 4 phase convs
    h' = conv3x3(a2) + s[y>>1, x>>1]                  conv2 + residual; epilogue also emits
    a' = act(bn1_next(h'))   (+ its 1-px frame)       ... the next conv's input, outer padding included
  SSM mode                       (StochasticSpatialModulation.forward, models/layers.py:228-234)
    taps = 3x3 tap stack of the noise map;  m1 = relu(conv1x1(taps))        (mlp_shared)
    a    = act((1+gamma) * bn0(x) + beta),  [gamma|beta] = conv3x3(m1)      (embed conv, modulation in its epilogue)

`halo points` mark the grid tensors that are inputs of a conv2d_lp: that is where the sequential
sub-image protocol (models/layers.py:103-143) and the row-band multi-GPU split patch the frame.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import packing as PK
from .config import GenConfig
from .ops import AttentionOp, ConvOp, Grid, SsmOp, c_store, pair_fold_eligible

PRECISIONS = {
    # name: (torch dtype, conv implementation)
    "fp32": (torch.float32, L.IMPL_SPLIT),      # exact mode (the <= 1e-3 gate) on tensor cores: fp32 tensors, operands as two fp16 terms, 3 MMAs per step
    "fp32-direct": (torch.float32, L.IMPL_DIRECT),   # exact mode on CUDA cores: cross-check of the split-precision kernel
    "fp16": (torch.float16, L.IMPL_AUTO),       # tcgen05 kind::f16, fp16 operands / fp32 accumulate
    "fp16-stream": (torch.float16, L.IMPL_UMMA),     # force the per-tap streaming kernel everywhere, SSM as two launches (A/B comparison)
    "fp16-direct": (torch.float16, L.IMPL_DIRECT),   # on-device cross-check of the tcgen05 kernel
}
# There is deliberately no 'bf16' entry.  Single-pass bf16 operands (7 mantissa bits) measure 2.7e-2 ... 6.8e-2 max-abs on the [-1, 1] image at
# random init (SURVEY 7.4 predicted 4e-2 ... 1e-1), above the <= 2e-2 gate of the 16-bit mode; fp16 operands run on the same tcgen05 kind::f16
# path at the same rate and meet it (<= 1.9e-2 on every fixture), and the Generator's activations stay below 14, far inside fp16's range.  The
# kernels keep their bf16 instantiations (launch-level tests compare them with the emulator), but the product does not offer the mode.

SSM_HIDDEN = 128   # nhidden of StochasticSpatialModulation (models/layers.py:220)
ALIGN = 1024       # arena alignment in bytes (TMA needs 16; keep tensors on separate 1 KiB lines)


def strip_module_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """test_sample.py:32-41 / utils.py:616-634: checkpoints saved from nn.DataParallel carry 'module.'."""
    return {(k.replace("module.", "") if "module" in k else k): v for k, v in sd.items()}


class PackedWeights:
    """All weights of one Generator in the operand layouts of libitg_b200 (packing.py), on `device`."""

    def __init__(self, cfg: GenConfig, sd: Dict[str, torch.Tensor], dtype: torch.dtype, device):
        sd = {k: v.detach().to("cpu") for k, v in strip_module_prefix(sd).items()}
        self.cfg, self.dtype = cfg, dtype
        t: Dict[str, torch.Tensor] = {}

        def put(name, v):
            t[name] = v.to(device).contiguous()

        def conv3(prefix, up=False):
            wgt = sd[prefix + "weight"]
            put(prefix + "w", PK.pack_upconv(wgt, dtype) if up else PK.pack_conv3x3(wgt, dtype))
            put(prefix + "b", PK.pad_vec(sd[prefix + "bias"], t[prefix + "w"].shape[1]))

        def conv1(prefix):
            put(prefix + "w", PK.pack_conv1x1(sd[prefix + "weight"], dtype))
            put(prefix + "b", PK.pad_vec(sd[prefix + "bias"], t[prefix + "w"].shape[1]))

        def bn(prefix, c):
            s, sh = PK.bn_fold(sd, prefix, max(PK.n_pad_of(c), c_store(c)))
            put(prefix + "scale", s)
            put(prefix + "shift", sh)

        def ssm(prefix, c):
            mean, rstd = PK.bn_stats(sd, prefix + "bn.", c_store(c))
            put(prefix + "mean", mean)
            put(prefix + "rstd", rstd)
            w1 = sd[prefix + "mlp_shared.0.weight"].float().reshape(1, SSM_HIDDEN, 9)      # (128,1,3,3) -> taps as K
            put(prefix + "mlp.w", PK._pad_nk(w1, SSM_HIDDEN, 16).to(dtype))
            put(prefix + "mlp.b", PK.pad_vec(sd[prefix + "mlp_shared.0.bias"], SSM_HIDDEN))
            if dtype != torch.float32:          # operand of the fused kernel (itg_ssm_fwd): taps + bias hi/lo in one K = 16 row
                put(prefix + "mlp.wf", PK.pack_ssm_mlp(sd[prefix + "mlp_shared.0.weight"], sd[prefix + "mlp_shared.0.bias"], dtype))
            we, be = PK.pack_ssm_embed(sd[prefix + "embed.weight"], sd[prefix + "embed.bias"], dtype)
            put(prefix + "embed.w", we)
            put(prefix + "embed.b", be)

        ssm_mode = cfg.type_norm == "SSM"
        conv3("start.conv.")
        for k, (ci, co) in enumerate(cfg.block_channels(), start=1):
            p = f"block{k}."
            conv3(p + "conv1.conv.", up=(k > 1 and not ssm_mode))
            conv3(p + "conv2.conv.")
            if ci != co:
                conv1(p + "conv3.")
            if ssm_mode and ci != co:              # bias of conv2 with the 1x1 shortcut folded in (one accumulator, itg_conv_desc.in2)
                put(p + "conv2f.b", t[p + "conv2.conv.b"] + t[p + "conv3.b"])
            if ssm_mode:
                ssm(p + "bn1.", ci)
                ssm(p + "bn2.", co)
                if ci != co:
                    ssm(p + "bn3.", ci)
            else:
                bn(p + "bn1.", ci)
                bn(p + "bn2.", co)
        if not ssm_mode:
            bn("bn.", cfg.final_ch)
        if cfg.attention:
            for n in ("theta", "phi", "g", "o"):
                wgt = sd[f"attention.{n}.weight"].float()
                put(f"attention.{n}.w", wgt.reshape(wgt.shape[0], wgt.shape[1]))
                put(f"attention.{n}.b", sd[f"attention.{n}.bias"].float())
            put("attention.gamma", sd["attention.gamma"].float().reshape(1))
        conv3("final.conv.")
        self.t = t

    def __getitem__(self, k):
        return self.t[k]

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.t.values())


@dataclass
class _VGrid:
    """Grid whose storage is assigned after the step list is known (liveness-based arena)."""
    name: str
    h: int
    w: int
    c: int
    dtype: torch.dtype
    first: int = -1
    last: int = -1
    grid: Optional[Grid] = None
    pinned: bool = False          # inputs / outputs that must not alias anything

    @property
    def nbytes(self) -> int:
        return (self.h + 2) * (self.w + 2) * self.c * torch.empty((), dtype=self.dtype).element_size()


@dataclass
class HaloPoint:
    """A conv2d_lp input: grid tensor `grid` feeds conv `name`; `r` = per-patch size at that tensor's resolution."""
    step: int            # index of the producing step: frame patching / pushes run right after it
    name: str
    grid: Grid
    r: int
    pull_step: int = -1  # index of the step right before the consumer conv (>= step): pulls may run as late as after it


class Plan:
    """The launch list of one Generator forward on a th x tw patch grid, with its buffers."""

    def __init__(self, cfg: GenConfig, weights: PackedWeights, backend, th: int, tw: int, device,
                 impl: int, img_layout: int = L.IMG_MERGED, reuse_buffers: bool = True, fuse_ssm: Optional[bool] = None,
                 pre_tanh: bool = False):
        if th < 1 or tw < 1:
            raise ValueError("patch grid must be at least 1 x 1")
        self.cfg, self.w, self.backend = cfg, weights, backend
        self.th, self.tw, self.device, self.impl = th, tw, device, impl
        self.dtype = weights.dtype
        # SSM norms as ONE launch each (itg_ssm_fwd, hidden map kept on chip) on the 16-bit tcgen05 path; the fp32 exact mode
        # and the forced-implementation precisions run mlp_shared and embed as two itg_conv_fwd launches
        if fuse_ssm is None:
            fuse_ssm = impl == L.IMPL_AUTO and self.dtype != torch.float32 and not os.environ.get("ITG_SSM_UNFUSED")
        self.fuse_ssm = bool(fuse_ssm) and cfg.type_norm == "SSM"
        # ... and the 1x1 shortcut of an SSM block (conv3 on the bn3 output, at the block's output resolution) folded into its conv2 as a second
        # input of the CTA-pair kernel where both tiles fit (blocks 4 and 5 of the 34 Generator): one launch and one tensor less per block
        self.fold_shortcut = self.fuse_ssm and not os.environ.get("ITG_NO_FOLD")
        self.pre_tanh = pre_tanh      # parity aid: the final conv also leaves its fp32 pre-activation (generators.py:119) in self.out_pre
        self.img_layout = img_layout
        self.border = L.BORDER_REPLICATE if cfg.border_is_replicate else L.BORDER_CONSTANT
        self._vgrids: List[_VGrid] = []
        self._steps: List[Tuple[str, object]] = []          # (kind, op builder args) before materialisation
        self.halo_points: List[HaloPoint] = []
        self._halo_tmp: List[Tuple[int, str, _VGrid, int]] = []
        b = cfg.base_res
        P = cfg.patch_px
        # static inputs / outputs (caller copies into / out of them; stable addresses for CUDA graphs)
        # (non-local Generator: the rings stay zero -- they ARE the zero padding of start / mlp_shared -- and set_inputs fills the interiors)
        alloc = torch.zeros if cfg.nonlocal_mode else torch.empty
        self.z_in = alloc((cfg.z_dim, th * b + 2, tw * b + 2), dtype=torch.float32, device=device)
        self.maps_in: List[torch.Tensor] = []
        if cfg.type_norm == "SSM":
            self.maps_in = [alloc((th * cfg.level_res(k) + 4, tw * cfg.level_res(k) + 4), dtype=torch.float32,
                                  device=device) for k in range(1, cfg.n_layers_G + 1)]
            if cfg.nonlocal_mode and not self.fuse_ssm:
                raise NotImplementedError("the non-local SSM Generator runs on the fused 16-bit kernel only (precision 'fp16')")
        if img_layout == L.IMG_MERGED:
            self.out = torch.empty((1, cfg.img_ch, th * P, tw * P), dtype=torch.float32, device=device)
        else:
            self.out = torch.empty((th * tw, cfg.img_ch, P, P), dtype=torch.float32, device=device)
        self.out_pre = torch.empty((th * P, tw * P, 8), dtype=torch.float32, device=device) if pre_tanh else None
        self._build()
        self._materialise(reuse_buffers)

    # ------------------------------------------------------------------------------------------
    # builder helpers
    # ------------------------------------------------------------------------------------------
    def _g(self, name: str, h: int, w: int, c: int) -> _VGrid:
        v = _VGrid(name, h, w, c, self.dtype)
        self._vgrids.append(v)
        return v

    def _touch(self, *vs):
        i = len(self._steps)
        for v in vs:
            if v is None:
                continue
            if v.first < 0:
                v.first = i
            v.last = i

    def _conv(self, name, mode, src: _VGrid, wkey: str, *, k=None, out_raw=None, out_act=None, norm=None,
              leak=None, linear=False, border=None, res=None, res_shift=0, mod=None,
              mod_shift=0, mod_prefix=None, window=None, out_hw=None, out_c=None, img=False, in2=None, w2key=None, bkey=None):
        self._touch(src, out_raw, out_act, res, mod, in2)
        self._note_consumer(src)
        self._steps.append(("conv", dict(name=name, mode=mode, src=src, wkey=wkey, k=k, out_raw=out_raw, out_act=out_act,
                                         norm=norm, leak=leak, linear=linear, border=border, res=res,
                                         res_shift=res_shift, mod=mod, mod_shift=mod_shift, mod_prefix=mod_prefix,
                                         window=window, out_hw=out_hw, out_c=out_c, img=img, in2=in2, w2key=w2key, bkey=bkey)))

    def _halo(self, name: str, g: _VGrid, r: int):
        self._halo_tmp.append([len(self._steps) - 1, name, g, r, -1])

    def _note_consumer(self, src: _VGrid):
        """Called before a conv reading `src` is appended: remember the last step before the consumer of a halo grid."""
        for h in self._halo_tmp:
            if h[2] is src and h[4] < 0:
                h[4] = len(self._steps) - 1

    # ------------------------------------------------------------------------------------------
    # network topology
    # ------------------------------------------------------------------------------------------
    def _build(self):
        cfg, th, tw = self.cfg, self.th, self.tw
        b = cfg.base_res
        H0, W0 = th * b, tw * b
        z = self._g("z", H0, W0, c_store(cfg.z_dim))
        self._touch(z)
        self._steps.append(("pack_z", z))
        if cfg.type_norm == "BN":
            a_last = self._build_bn(z)
        else:
            a_last = self._build_ssm(z)
        self._touch(a_last)
        self._note_consumer(a_last)
        self._steps.append(("conv", dict(name="final", mode=L.CONV3X3, src=a_last, wkey="final.conv.", k=None, out_raw=None,
                                         out_act=None, norm=None, leak=None, linear=False, border=None, res=None,
                                         res_shift=0, mod=None, mod_shift=0, mod_prefix=None, window=None, out_hw=None,
                                         out_c=None, img=True)))

    def _attention(self, x: _VGrid, out_raw: Optional[_VGrid], out_act: Optional[_VGrid], norm: Optional[str]):
        if self.cfg.nonlocal_mode and not (x.h == x.w and x.h in (8, 16)):
            raise NotImplementedError(
                f"non-local Generator with attention over a {x.h}x{x.w} feature map: the attention kernel serves one 8x8 or 16x16 map "
                "(z of 2x2 or 4x4, the sizes sample_from_gen(base_res=4) uses); larger maps need attention=False")
        self._touch(x, out_raw, out_act)
        self._steps.append(("att", dict(x=x, out_raw=out_raw, out_act=out_act, norm=norm)))

    def _build_bn(self, z: _VGrid) -> _VGrid:
        cfg, th, tw = self.cfg, self.th, self.tw
        chans = cfg.block_channels()
        n = cfg.n_layers_G
        H, W = z.h, z.w
        c0 = chans[0][0]
        h_raw = self._g("h0", H, W, c_store(c0))
        a = self._g("a.block1.conv1", H, W, c_store(c0))
        self._conv("start", L.CONV3X3, z, "start.conv.", out_raw=h_raw, out_act=a, norm="block1.bn1.", border=self.border)
        self._halo("block1.conv1", a, cfg.level_res(1))
        for k, (ci, co) in enumerate(chans, start=1):
            r = cfg.level_res(k)
            H, W = th * r, tw * r
            last = k == n
            p = f"block{k}."
            a2 = self._g(f"a.{p}conv2", H, W, c_store(co))
            if k == 1:
                self._conv(p + "conv1", L.CONV3X3, a, p + "conv1.conv.", out_act=a2, norm=p + "bn2.", border=self.border)
                res, res_shift = h_raw, 0
            else:
                self._conv(p + "conv1", L.UPCONV, a, p + "conv1.conv.", out_act=a2, norm=p + "bn2.", border=self.border,
                           out_hw=(H, W))
            self._halo(p + "conv2", a2, r)
            if k > 1:
                # the 1x1 shortcut does not need a2: it sits between conv1 and conv2 so that, in the row-band multi-GPU
                # split, a2's halo rows travel while it runs
                if ci != co:
                    s = self._g(f"s.{p}", H // 2, W // 2, c_store(co))
                    self._conv(p + "conv3", L.CONV1X1, h_raw, p + "conv3.", out_raw=s)
                    res = s
                else:
                    res = h_raw
                res_shift = 1
            att_here = k == 3 and cfg.attention
            next_norm = "bn." if last else f"block{k + 1}.bn1."
            next_name = "final" if last else f"block{k + 1}.conv1"
            a_new = self._g(f"a.{next_name}", H, W, c_store(co))
            if att_here:
                h_mid = self._g(f"h{k}.pre_att", H, W, c_store(co))
                self._conv(p + "conv2", L.CONV3X3, a2, p + "conv2.conv.", out_raw=h_mid, res=res, res_shift=res_shift)
                h_new = None if last else self._g(f"h{k}", H, W, c_store(co))
                self._attention(h_mid, h_new, a_new, next_norm)
            else:
                h_new = None if last else self._g(f"h{k}", H, W, c_store(co))
                self._conv(p + "conv2", L.CONV3X3, a2, p + "conv2.conv.", out_raw=h_new, out_act=a_new, norm=next_norm,
                           border=self.border, res=res, res_shift=res_shift)
            self._halo(next_name, a_new, r)
            h_raw, a = h_new, a_new
        return a

    def _ssm_norm(self, prefix: str, taps: Optional[_VGrid], x: _VGrid, x_shift: int, out: _VGrid, linear: bool, border: int,
                  level: int = 0):
        """out = [act]((1 + gamma) * bn0(x) + beta) with [gamma|beta] = embed(relu(mlp_shared(map)))."""
        H, W = out.h, out.w
        if self.fuse_ssm:
            self._touch(x, out)
            self._steps.append(("ssm", dict(prefix=prefix, level=level, x=x, x_shift=x_shift, out=out, linear=linear, border=border)))
            return
        m1 = self._g(f"m1.{prefix}", H + 2, W + 2, SSM_HIDDEN)
        self._conv(prefix + "mlp", L.CONV1X1, taps, prefix + "mlp.", k=16, out_act=m1, leak=0.0, border=L.BORDER_NONE)
        self._conv(prefix + "embed", L.CONV3X3, m1, prefix + "embed.", k=SSM_HIDDEN, out_act=out, linear=linear,
                   border=border, mod=x, mod_shift=x_shift, mod_prefix=prefix, window=(H, W), out_c=out.c)

    def _build_ssm(self, z: _VGrid) -> _VGrid:
        cfg, th, tw = self.cfg, self.th, self.tw
        chans = cfg.block_channels()
        n = cfg.n_layers_G
        c0 = chans[0][0]
        h_raw = self._g("h0", z.h, z.w, c_store(c0))
        self._conv("start", L.CONV3X3, z, "start.conv.", out_raw=h_raw)
        a_last = None
        for k, (ci, co) in enumerate(chans, start=1):
            r = cfg.level_res(k)
            H, W = th * r, tw * r
            last = k == n
            p = f"block{k}."
            x_shift = 0 if k == 1 else 1
            taps = None
            if not self.fuse_ssm:
                taps = self._g(f"taps{k}", H + 2, W + 2, 16)
                self._touch(taps)
                self._steps.append(("pack_map", (k - 1, taps)))
            a1 = self._g(f"a.{p}conv1", H, W, c_store(ci))
            self._ssm_norm(p + "bn1.", taps, h_raw, x_shift, a1, False, self.border, k - 1)
            self._halo(p + "conv1", a1, r)
            t = self._g(f"t.{p}", H, W, c_store(co))
            self._conv(p + "conv1", L.CONV3X3, a1, p + "conv1.conv.", out_raw=t)
            a2 = self._g(f"a.{p}conv2", H, W, c_store(co))
            self._ssm_norm(p + "bn2.", taps, t, 0, a2, False, self.border, k - 1)
            self._halo(p + "conv2", a2, r)
            fold = {}
            if ci != co:
                s_in = self._g(f"sin.{p}", H, W, c_store(ci))
                self._ssm_norm(p + "bn3.", taps, h_raw, x_shift, s_in, True, L.BORDER_NONE, k - 1)
                if self.fold_shortcut and pair_fold_eligible(c_store(co), c_store(ci), co):
                    fold = dict(in2=s_in, w2key=p + "conv3.", bkey=p + "conv2f.b")     # conv2 = conv3x3(a2) + conv1x1(s_in), one launch
                    res, res_shift = None, 0
                else:
                    s = self._g(f"s.{p}", H, W, c_store(co))
                    self._conv(p + "conv3", L.CONV1X1, s_in, p + "conv3.", out_raw=s)
                    res, res_shift = s, 0
            else:
                res, res_shift = h_raw, x_shift
            att_here = k == 3 and cfg.attention
            if last and not att_here:
                a_last = self._g("a.final", H, W, c_store(co))
                self._conv(p + "conv2", L.CONV3X3, a2, p + "conv2.conv.", out_act=a_last, border=self.border, res=res,
                           res_shift=res_shift, **fold)
                h_raw = None
            else:
                h_new = self._g(f"h{k}", H, W, c_store(co))
                self._conv(p + "conv2", L.CONV3X3, a2, p + "conv2.conv.", out_raw=h_new, res=res, res_shift=res_shift, **fold)
                if att_here:
                    h_att = self._g(f"h{k}.att", H, W, c_store(co))
                    self._attention(h_new, h_att, None, None)
                    h_new = h_att
                h_raw = h_new
        self._halo("final", a_last, cfg.level_res(n))
        return a_last

    # ------------------------------------------------------------------------------------------
    # buffers and compiled launch list
    # ------------------------------------------------------------------------------------------
    def _materialise(self, reuse: bool):
        # liveness-based arena: tensors whose lifetimes [first, last] do not overlap may share bytes.  Offsets are
        # assigned greedily by decreasing size (the large full-resolution tensors come last in the network, so a
        # first-come first-fit wastes most of the reuse): each tensor takes the lowest offset that does not collide
        # with an already placed tensor alive at the same time.
        offs: Dict[int, int] = {}
        total = 0
        if reuse:
            placed: List[Tuple[int, int, int, int]] = []          # (offset, size, first, last)
            for v in sorted(self._vgrids, key=lambda v: -v.nbytes):
                need = (v.nbytes + ALIGN - 1) // ALIGN * ALIGN
                busy = sorted((o, o + sz) for o, sz, f, l in placed if not (l < v.first or f > v.last))
                off = 0
                for lo, hi in busy:
                    if off + need <= lo:
                        break
                    off = max(off, hi)
                placed.append((off, need, v.first, v.last))
                offs[id(v)] = off
                total = max(total, off + need)
        else:
            for v in self._vgrids:
                need = (v.nbytes + ALIGN - 1) // ALIGN * ALIGN
                offs[id(v)] = total
                total += need
        self.arena = torch.empty(total + ALIGN, dtype=torch.uint8, device=self.device)
        base_off = (-self.arena.data_ptr()) % ALIGN if self.arena.is_cuda else 0
        self.arena_bytes = total
        es = torch.empty((), dtype=self.dtype).element_size()
        for v in self._vgrids:
            o = base_off + offs[id(v)]
            n = (v.h + 2) * (v.w + 2) * v.c
            buf = self.arena[o:o + n * es].view(self.dtype).view(v.h + 2, v.w + 2, v.c)
            v.grid = Grid(buf, v.h, v.w, v.c)
        self.grids = {v.name: v.grid for v in self._vgrids}
        self.halo_points = [HaloPoint(step, name, g.grid, r, max(pull, step)) for step, name, g, r, pull in self._halo_tmp]

        be, w = self.backend, self.w
        self.ops: List[Tuple[str, object]] = []
        self.fns: List[Callable[[], None]] = []
        for kind, a in self._steps:
            if kind == "pack_z":
                zg = a.grid
                self.ops.append(("pack_z", zg))
                self.fns.append(be.compile_pack_nchw(self.z_in, zg.buf))
            elif kind == "pack_map":
                lvl, taps = a
                self.ops.append(("pack_map", taps.grid))
                self.fns.append(be.compile_pack_map_taps(self.maps_in[lvl], taps.grid))
            elif kind == "att":
                op = self._attention_op(a)
                self.ops.append(("att", op))
                self.fns.append(be.compile_attention(op))
            elif kind == "ssm":
                pre = a["prefix"]
                op = SsmOp(map=self.maps_in[a["level"]], w_mlp=w[pre + "mlp.wf"], w_embed=w[pre + "embed.w"], b_embed=w[pre + "embed.b"],
                           x=a["x"].grid, x_shift=a["x_shift"], mean=w[pre + "mean"], rstd=w[pre + "rstd"], out=a["out"].grid,
                           leak=self.cfg.leak, linear=a["linear"], border=a["border"], zero_ring=self.cfg.nonlocal_mode, name=pre + "ssm")
                self.ops.append(("ssm", op))
                self.fns.append(be.compile_ssm(op))
            else:
                op = self._conv_op(a)
                self.ops.append(("conv", op))
                self.fns.append(be.compile_conv(op))
        del self._steps, self._halo_tmp

    def _norm_vectors(self, norm: Optional[str]):
        if norm is None:
            return None, None
        return self.w[norm + "scale"], self.w[norm + "shift"]

    def _conv_op(self, a: dict) -> ConvOp:
        cfg, w = self.cfg, self.w
        src: Grid = a["src"].grid
        wt = w[a["wkey"] + "w"]
        op = ConvOp(mode=a["mode"], src=src, w=wt, k=a["k"] if a["k"] is not None else src.c,
                    bias=w[a["bkey"]] if a.get("bkey") else w[a["wkey"] + "b"], impl=self.impl, name=a["name"])
        if a.get("in2") is not None:                # folded 1x1 shortcut (CTA-pair kernel)
            op.in2, op.w2, op.k2, op.impl = a["in2"].grid, w[a["w2key"] + "w"], a["in2"].grid.c, L.IMPL_PAIR
        if a["window"] is not None:                 # valid conv on the over-sized SSM hidden map
            H, W = a["window"]
            op.in_h, op.in_w, op.in_pitch = H, W, src.w + 2
            op.in_elem_off = ((src.w + 2) + 1) * src.c
        mh, mw = op.m_h, op.m_w
        op.out_h, op.out_w = a["out_hw"] if a["out_hw"] is not None else (mh, mw)
        if a["img"]:
            op.out_c = 8
            if self.pre_tanh:
                op.out_f32 = self.out_pre
                return op
            op.out_img, op.img_c, op.img_layout, op.patch = self.out, cfg.img_ch, self.img_layout, cfg.patch_px
            return op
        outs = [g for g in (a["out_raw"], a["out_act"]) if g is not None]
        op.out_c = a["out_c"] if a["out_c"] is not None else outs[0].c
        op.out_raw = a["out_raw"].grid if a["out_raw"] is not None else None
        op.out_act = a["out_act"].grid if a["out_act"] is not None else None
        op.scale, op.shift = self._norm_vectors(a["norm"])
        op.leak = cfg.leak if a["leak"] is None else a["leak"]
        op.act_linear = a["linear"]
        op.border = a["border"] if a["border"] is not None else L.BORDER_NONE
        if a["res"] is not None:
            r: Grid = a["res"].grid
            op.res_kind, op.res, op.res_shift = L.RES_GRID, r.buf, a["res_shift"]
            op.res_c, op.res_h, op.res_w = r.c, r.h, r.w
        if a["mod"] is not None:
            op.mod_x, op.mod_shift = a["mod"].grid, a["mod_shift"]
            op.mod_mean, op.mod_rstd = w[a["mod_prefix"] + "mean"], w[a["mod_prefix"] + "rstd"]
        return op

    def _attention_op(self, a: dict) -> AttentionOp:
        cfg, w = self.cfg, self.w
        scale, shift = self._norm_vectors(a["norm"])
        th, tw, patch = self.th, self.tw, cfg.level_res(3)
        if cfg.nonlocal_mode:                      # the non-local Generator attends over the WHOLE feature map (layers.py:246-258 on the full image)
            th, tw, patch = 1, 1, a["x"].h
        return AttentionOp(
            x=a["x"].grid, th=th, tw=tw, patch=patch, C=2 * cfg.G_ch,
            w_theta=w["attention.theta.w"], b_theta=w["attention.theta.b"], w_phi=w["attention.phi.w"],
            b_phi=w["attention.phi.b"], w_g=w["attention.g.w"], b_g=w["attention.g.b"], w_o=w["attention.o.w"],
            b_o=w["attention.o.b"], gamma=w["attention.gamma"],
            out_raw=a["out_raw"].grid if a["out_raw"] is not None else None,
            out_act=a["out_act"].grid if a["out_act"] is not None else None,
            scale=scale, shift=shift, leak=cfg.leak, border=self.border)

    # ------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------
    @property
    def n_launches(self) -> int:
        return len(self.fns)

    def run(self, hooks: Optional[Dict[int, Callable[[], None]]] = None) -> torch.Tensor:
        """Launch every step on the current stream; `hooks[i]` runs right after step i (halo patching)."""
        if hooks:
            for i, fn in enumerate(self.fns):
                fn()
                h = hooks.get(i)
                if h is not None:
                    h()
        else:
            for fn in self.fns:
                fn()
        return self.out

    def set_inputs(self, z: torch.Tensor, maps: Optional[Sequence[torch.Tensor]] = None) -> None:
        """Copy host- or device-resident fp32 noise into the plan's static input buffers (async on the current stream)."""
        if self.cfg.nonlocal_mode:
            return self._set_inputs_nonlocal(z, maps)
        z = z.reshape(self.z_in.shape) if z.dim() == 4 else z
        if tuple(z.shape) != tuple(self.z_in.shape):
            raise ValueError(f"z has shape {tuple(z.shape)}, the {self.th}x{self.tw} patch grid needs {tuple(self.z_in.shape)}")
        self.z_in.copy_(z, non_blocking=True)
        if self.cfg.type_norm == "SSM":
            if maps is None or len(maps) < len(self.maps_in):
                raise ValueError("SSM Generator needs one noise map per level (utils.py:237-256)")
            for dst, m in zip(self.maps_in, maps):
                m = m.reshape(dst.shape) if m.dim() == 4 else m
                if tuple(m.shape) != tuple(dst.shape):
                    raise ValueError(f"map has shape {tuple(m.shape)}, expected {tuple(dst.shape)}")
                dst.copy_(m, non_blocking=True)


    def _set_inputs_nonlocal(self, z: torch.Tensor, maps: Optional[Sequence[torch.Tensor]]) -> None:
        """Non-local Generator (sample_from_gen, utils.py:530-575): z is the bare (z_dim, h, w) latent and the maps are r x r; the zero rings
        of the static buffers are the zero padding of the first convs."""
        z = z[0] if z.dim() == 4 else z
        want = (self.cfg.z_dim, self.z_in.shape[1] - 2, self.z_in.shape[2] - 2)
        if tuple(z.shape) != want:
            raise ValueError(f"z has shape {tuple(z.shape)}, this plan needs {want}")
        self.z_in[:, 1:-1, 1:-1].copy_(z, non_blocking=True)
        if self.cfg.type_norm == "SSM":
            if maps is None or len(maps) < len(self.maps_in):
                raise ValueError("SSM Generator needs one noise map per level (utils.py:558-564)")
            for dst, m in zip(self.maps_in, maps):
                m = m.reshape(m.shape[-2], m.shape[-1])
                if (m.shape[0] + 4, m.shape[1] + 4) != tuple(dst.shape):
                    raise ValueError(f"map has shape {tuple(m.shape)}, expected {(dst.shape[0] - 4, dst.shape[1] - 4)}")
                dst[2:-2, 2:-2].copy_(m, non_blocking=True)


class Engine:
    """Packed weights + cached plans (+ optional CUDA graphs) of one Generator."""

    def __init__(self, cfg: GenConfig, state_dict: Dict[str, torch.Tensor], precision: str = "fp16", device="cuda",
                 backend=None):
        if precision not in PRECISIONS:
            hint = (" (single-pass bf16 operands cannot meet the 2e-2 image tolerance; 'fp16' runs on the same tensor-core path at the same rate)"
                    if precision.startswith("bf16") else "")
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}{hint}")
        if cfg.nonlocal_mode:
            # the non-local Generator is fully convolutional over the whole image: plans address the h x w latent grid as an h x w grid
            # of 1-pixel "patches" (base_res 1) with a zero frame
            import dataclasses
            cfg = dataclasses.replace(cfg, base_res=1)
        self.cfg, self.precision = cfg, precision
        self.dtype, self.impl = PRECISIONS[precision]
        self.device = torch.device(device)
        if backend is None:
            from .ops import CudaBackend
            if self.device.type != "cuda":
                raise L.ItgError("infinite_texture_gans_b200 computes on CUDA devices only (there is no CPU fallback)")
            backend = CudaBackend()
        self.backend = backend
        self.weights = PackedWeights(cfg, state_dict, self.dtype, self.device)
        self._plans: Dict[Tuple[int, int, int, bool], Plan] = {}
        self._graphs: Dict[Tuple[int, int, int], "torch.cuda.CUDAGraph"] = {}

    def plan(self, th: int, tw: int, img_layout: int = L.IMG_MERGED, pre_tanh: bool = False) -> Plan:
        key = (th, tw, img_layout, pre_tanh)
        p = self._plans.get(key)
        if p is None:
            p = Plan(self.cfg, self.weights, self.backend, th, tw, self.device, self.impl, img_layout, pre_tanh=pre_tanh)
            self._plans[key] = p
        return p

    def forward_pre_tanh(self, z: torch.Tensor, maps: Optional[Sequence[torch.Tensor]] = None, *, th: int, tw: int) -> torch.Tensor:
        """Parity aid: the final conv's fp32 output BEFORE the tanh (generators.py:119-121) as (1, img_ch, th*P, tw*P)."""
        with self._on_device():
            p = self.plan(th, tw, L.IMG_MERGED, pre_tanh=True)
            p.set_inputs(z, maps)
            p.run()
        return p.out_pre[..., : self.cfg.img_ch].permute(2, 0, 1).unsqueeze(0)

    def drop_plans(self) -> None:
        self._plans.clear()
        self._graphs.clear()

    def forward(self, z: torch.Tensor, maps: Optional[Sequence[torch.Tensor]] = None, *, th: int, tw: int,
                img_layout: int = L.IMG_MERGED, graph: bool = False) -> torch.Tensor:
        """One-shot Generator forward of the whole th x tw grid.  Returns the plan's output buffer
        ((1, img_ch, th*P, tw*P) merged or (th*tw, img_ch, P, P) patches); it is overwritten by the next call."""
        with self._on_device():
            p = self.plan(th, tw, img_layout)
            p.set_inputs(z, maps)
            if graph and self.device.type == "cuda":
                self.replay(th, tw, img_layout)
            else:
                p.run()
        return p.out

    def _on_device(self):
        """Launches go to the current stream of the CURRENT device: make the engine's device current (one process may drive several GPUs)."""
        import contextlib
        return torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()

    def replay(self, th: int, tw: int, img_layout: int = L.IMG_MERGED) -> torch.Tensor:
        """Run the plan's launch list from a captured CUDA graph (inputs already in plan.z_in / plan.maps_in)."""
        key = (th, tw, img_layout)
        with self._on_device():
            p = self.plan(th, tw, img_layout)
            g = self._graphs.get(key)
            if g is None:
                p.run()                              # warm-up outside capture: one-time attribute / tensor-map setup
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    p.run()
                self._graphs[key] = g
            g.replay()
        return p.out
