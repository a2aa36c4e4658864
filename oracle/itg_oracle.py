"""CPU oracle for the patch-by-patch Generator inference path (TEST INFRASTRUCTURE ONLY).

This module is a from-scratch CPU restatement (torch CPU fp32/fp64 functional ops, no nn.Module, no
CUDA) of the reference algorithm for the path

    test_sample.py -> utils.sample_from_gen_PatchByPatch_test -> ResidualPatchGenerator.forward
    (--padding_mode local)

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product (``infinite_texture_gans_b200``) never does.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the *unmodified* reference imported from /root/reference in the build
container; those outputs are committed under ``tests/golden/`` together with the generating script
``tests/golden/make_golden.py`` and replayed by ``tests/test_oracle_golden.py``.

All ``file:line`` citations are relative to the reference repository root.

Two formulations are provided and checked against each other and against the reference:

* ``forward_patches`` -- literal restatement: patches are merged, padded (outer padding or stored
  halo state), re-cropped with a 1-px overlap and pushed through *valid* 3x3 convs, with the
  sequential halo protocol of ``LocalPadder`` (models/layers.py:78-173).  Driven by
  ``sample_patch_by_patch`` it reproduces ``sample_from_gen_PatchByPatch_test`` (utils.py:258-397),
  i.e. "Oracle B" of SURVEY.md section 3.4.
* ``forward_merged`` -- the one-shot formulation on the merged grid ("Oracle A"): every local-padding
  conv is a 3x3 'same' conv of the whole merged image with replicate/constant border; attention is
  evaluated per patch.  This is what the CUDA path computes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5  # nn.BatchNorm2d default, models/layers.py:279-280, models/generators.py:78


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class GenCfg:
    """Constructor arguments of ResidualPatchGenerator (models/generators.py:25-28)."""
    z_dim: int = 128
    G_ch: int = 64
    base_res: int = 4
    n_layers_G: int = 4
    attention: bool = True
    img_ch: int = 3
    leak: float = 0.0
    type_norm: str = "BN"
    map_dim: int = 1
    outer_padding: str = "replicate"   # F.pad mode: 'replicate' | 'constant'
    num_patches_h: int = 3
    num_patches_w: int = 3

    @property
    def patch_px(self) -> int:                      # utils.py:294
        return (2 ** (self.n_layers_G - 1)) * self.base_res

    def block_channels(self) -> List[Tuple[int, int]]:
        """(in, out) channels of block1..block_n (models/generators.py:61-74)."""
        c = self.G_ch
        chans = [(8 * c, 8 * c), (8 * c, 4 * c), (4 * c, 2 * c), (2 * c, c)]
        if self.n_layers_G >= 5:
            chans.append((c, c // 2))
            if self.n_layers_G == 6:
                chans.append((c // 2, c // 4))
        return chans

    @property
    def final_ch(self) -> int:
        return self.block_channels()[-1][1]


def _pad_mode(outer_padding: str) -> str:
    # 'zeros' is not an F.pad mode in the reference; BASELINE's "zeros" means 'constant'.
    return "constant" if outer_padding in ("zeros", "constant") else outer_padding


# --------------------------------------------------------------------------------------------
# geometry and patch <-> image plumbing
# --------------------------------------------------------------------------------------------
def geometry(out_h: int, out_w: int, cfg: GenCfg) -> Dict[str, int]:
    """Step / patch-count arithmetic of the test sampler (utils.py:294-303)."""
    P = cfg.patch_px
    steps_h = math.ceil((out_h / P - 1) / (cfg.num_patches_h - 1))
    steps_w = math.ceil((out_w / P - 1) / (cfg.num_patches_w - 1))
    return dict(P=P, steps_h=steps_h, steps_w=steps_w,
                total_h=steps_h * (cfg.num_patches_h - 1) + 1,
                total_w=steps_w * (cfg.num_patches_w - 1) + 1)


def crop_windows(img: Tensor, win_h: int, win_w: int, stride: int) -> Tensor:
    """Sliding-window crop, row-major window order, images outermost (utils.py:658-742).

    (N,C,H,W) -> (N*P,C,win_h,win_w).  The reference uses ONE stride for both axes.
    Always returns float32, like the reference (its torch.cat seed is an fp32 empty tensor,
    utils.py:681,714) unless the input is float64 (type promotion)."""
    N, C, H, W = img.shape
    ys = range(0, H - win_h + 1, stride)
    xs = range(0, W - win_w + 1, stride)
    out = [img[n, :, y:y + win_h, x:x + win_w] for n in range(N) for y in ys for x in xs]
    res = torch.stack(out, 0)
    return res if res.dtype == torch.float64 else res.float()


def merge_patches(patches: Tensor, rows: int, cols: int) -> Tensor:
    """Inverse of a non-overlapping crop (utils.py:577-613): (B,C,h,w) -> (B/(rows*cols),C,rows*h,cols*w)."""
    B, C, h, w = patches.shape
    n = B // (rows * cols)
    x = patches.reshape(n, rows, cols, C, h, w).permute(0, 3, 1, 4, 2, 5)
    res = x.reshape(n, C, rows * h, cols * w)
    return res if res.dtype == torch.float64 else res.float()


# --------------------------------------------------------------------------------------------
# LocalPadder (models/layers.py:38-173)
# --------------------------------------------------------------------------------------------
@dataclass
class PadState:
    """Eval-mode halo state of ONE LocalPadder instance (models/layers.py:69-76)."""
    col: Optional[Tensor] = None            # vertical_padding_variable
    col_next: Optional[Tensor] = None       # vertical_padding_variable_next_image
    row: Optional[Tensor] = None            # horizontal_padding_variable
    row_cur: Optional[Tensor] = None        # horizontal_padding_variable_for_current_row
    row_next: Optional[Tensor] = None       # horizontal_padding_variable_for_next_row


def _update_state(st: PadState, merged: Tensor, loc: str, h: int, w: int, cfg: GenCfg) -> None:
    """models/layers.py:103-143.  h,w = per-patch size of the merged input."""
    nph, npw = cfg.num_patches_h, cfg.num_patches_w
    mode = _pad_mode(cfg.outer_padding)
    if st.col_next is not None:                                            # :105-106
        st.col = st.col_next
    last_col = "last_col" in loc
    st.col_next = None if last_col else merged[:, :, :, [w * (npw - 1) - 1]]   # :110-113
    r = h * (nph - 1) - 1
    row_slice = merged[:, :, [r], :] if last_col else merged[:, :, [r], : w * (npw - 1)]  # :115-120
    if "1st_col" in loc:                                                   # :122-128
        if "1st_row" not in loc:
            st.row_cur = F.pad(st.row_next.clone(), (1, 1, 0, 0), mode)
        st.row_next = row_slice
    else:                                                                  # :129-132
        st.row_next = torch.cat((st.row_next, row_slice), -1)
    if st.row_cur is not None:                                             # :138-143
        st.row = st.row_cur[:, :, :, : npw * w + 2].clone()
        st.row_cur = None if last_col else st.row_cur[:, :, :, (npw - 1) * w:]


def _pad_merged(st: PadState, merged: Tensor, loc: str, cfg: GenCfg, training: bool) -> Tensor:
    """models/layers.py:78-101: outer padding and/or stored halos -> (.,.,H+2,W+2)."""
    mode = _pad_mode(cfg.outer_padding)
    first_row, first_col = "1st_row" in loc, "1st_col" in loc
    if training or (first_row and first_col):
        return F.pad(merged, (1, 1, 1, 1), mode)
    if first_row:
        out = torch.cat((st.col, merged), -1)
        return F.pad(out, (0, 1, 1, 1), mode)
    if first_col:
        out = F.pad(merged, (1, 1, 0, 1), mode)
        return torch.cat((st.row, out), -2)
    out = torch.cat((st.col, merged), -1)
    out = F.pad(out, (0, 1, 0, 1), mode)
    return torch.cat((st.row, out), -2)


def local_pad(x: Tensor, loc: str, st: PadState, cfg: GenCfg, *, merge: bool = True,
              training: bool = False) -> Tensor:
    """LocalPadder.forward (models/layers.py:145-173): (B,C,r,r) -> (B,C,r+2,r+2).

    merge=False is the `start` layer: x is the already merged, randomly bordered z grid and is only
    re-cropped (models/layers.py:152-155,165-166)."""
    nph, npw = cfg.num_patches_h, cfg.num_patches_w
    _, _, H, W = x.shape
    if merge:
        merged = merge_patches(x, nph, npw)
        h, w = H, W
    else:
        h, w = H // nph, W // npw
        merged = x
    if not training:
        _update_state(st, merged, loc, h, w, cfg)
    if merge:
        merged = _pad_merged(st, merged, loc, cfg, training)
    return crop_windows(merged, w + 2, w + 2, w)


# --------------------------------------------------------------------------------------------
# elementwise pieces
# --------------------------------------------------------------------------------------------
def _act(x: Tensor, leak: float) -> Tensor:
    # models/generators.py:54-57, models/layers.py:289-292
    return F.leaky_relu(x, leak) if leak > 0 else F.relu(x)


def _bn_eval(x: Tensor, sd: Dict[str, Tensor], prefix: str, affine: bool = True) -> Tensor:
    """nn.BatchNorm2d in eval mode (running statistics, eps 1e-5)."""
    mean = sd[prefix + "running_mean"].to(x.dtype)
    var = sd[prefix + "running_var"].to(x.dtype)
    w = sd[prefix + "weight"].to(x.dtype) if affine else None
    b = sd[prefix + "bias"].to(x.dtype) if affine else None
    return F.batch_norm(x, mean, var, w, b, False, 0.0, BN_EPS)


def _conv(x: Tensor, sd: Dict[str, Tensor], prefix: str, padding: int = 0) -> Tensor:
    return F.conv2d(x, sd[prefix + "weight"].to(x.dtype), sd[prefix + "bias"].to(x.dtype), padding=padding)


def ssm(x: Tensor, m: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """StochasticSpatialModulation.forward (models/layers.py:228-234), local-padding variant (p=0):
    x (B,C,r,r), m (B,1,r+4,r+4) -> (1+gamma)*BN0(x)+beta."""
    out = _bn_eval(x, sd, prefix + "bn.", affine=False)
    actv = F.relu(_conv(m.to(x.dtype), sd, prefix + "mlp_shared.0."))
    emb = _conv(actv, sd, prefix + "embed.")
    gamma, beta = emb.chunk(2, dim=1)
    return (1 + gamma) * out + beta


def attention(x: Tensor, sd: Dict[str, Tensor], prefix: str = "attention.") -> Tensor:
    """Attention.forward (models/layers.py:246-258).  Per batch element (= per patch); no 1/sqrt(d)."""
    b, c, h, w = x.shape
    theta = _conv(x, sd, prefix + "theta.")
    phi = F.max_pool2d(_conv(x, sd, prefix + "phi."), [2, 2])
    g = F.max_pool2d(_conv(x, sd, prefix + "g."), [2, 2])
    theta = theta.reshape(b, c // 8, -1)
    phi = phi.reshape(b, c // 8, -1)
    g = g.reshape(b, c // 2, -1)
    beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
    o = torch.bmm(g, beta.transpose(1, 2)).reshape(b, c // 2, h, w)
    o = _conv(o, sd, prefix + "o.")
    return sd[prefix + "gamma"].to(x.dtype) * o + x


# --------------------------------------------------------------------------------------------
# Formulation 1: literal patch protocol (Oracle B building block)
# --------------------------------------------------------------------------------------------
@dataclass
class SeqState:
    """One PadState per conv2d_lp instance, keyed by its state-dict prefix."""
    pads: Dict[str, PadState] = field(default_factory=dict)

    def get(self, name: str) -> PadState:
        return self.pads.setdefault(name, PadState())


def _conv_lp(x: Tensor, sd, prefix: str, loc: str, state: SeqState, cfg: GenCfg, *, merge=True,
             training=False) -> Tensor:
    """conv2d_lp.forward (models/layers.py:29-36): LocalPadder then valid 3x3 conv."""
    xp = local_pad(x, loc, state.get(prefix), cfg, merge=merge, training=training)
    return _conv(xp.to(x.dtype), sd, prefix + "conv.")


def _norm(x, m, sd, prefix, cfg):
    if cfg.type_norm == "SSM":
        return ssm(x, m, sd, prefix)
    return _bn_eval(x, sd, prefix)


def _res_block(x, m, sd, prefix, loc, state, cfg: GenCfg, cin, cout, training=False):
    """ResBlockGenerator.forward (models/layers.py:301-322)."""
    out = _act(_norm(x, m, sd, prefix + "bn1.", cfg), cfg.leak)
    out = _conv_lp(out, sd, prefix + "conv1.", loc, state, cfg, training=training)
    out = _act(_norm(out, m, sd, prefix + "bn2.", cfg), cfg.leak)
    out = _conv_lp(out, sd, prefix + "conv2.", loc, state, cfg, training=training)
    sc = x
    if cin != cout:                                          # learnable shortcut, :294-299
        if cfg.type_norm == "SSM":
            sc = ssm(sc, m, sd, prefix + "bn3.")
        sc = _conv(sc, sd, prefix + "conv3.")
    return out + sc


def forward_patches(sd: Dict[str, Tensor], cfg: GenCfg, z: Tensor, maps: Optional[Sequence[Tensor]] = None,
                    image_location: str = "1st_row_1st_col", state: Optional[SeqState] = None,
                    training: bool = False, pre_tanh: bool = False) -> Tensor:
    """ResidualPatchGenerator.forward (models/generators.py:86-124) on patch batches.

    z: (N,z_dim,nph*base+2,npw*base+2) merged grid with its random 1-px ring.
    maps: per level i, (N*nph*npw, 1, r_i+4, r_i+4) patch crops (SSM only).
    Returns (N*nph*npw, img_ch, P, P)."""
    state = state if state is not None else SeqState()
    maps = list(maps) if maps is not None else [None] * cfg.n_layers_G
    loc = image_location
    h = _conv_lp(z, sd, "start.", loc, state, cfg, merge=False, training=training)
    for k, (cin, cout) in enumerate(cfg.block_channels(), start=1):
        if k > 1:
            h = F.interpolate(h, scale_factor=2, mode="nearest")          # generators.py:52,95-111
        h = _res_block(h, maps[k - 1], sd, f"block{k}.", loc, state, cfg, cin, cout, training)
        if k == 3 and cfg.attention:                                      # generators.py:101-102
            h = attention(h, sd)
    if cfg.type_norm == "BN":
        h = _bn_eval(h, sd, "bn.")
    h = _act(h, cfg.leak)
    h = _conv_lp(h, sd, "final.", loc, state, cfg, training=training)
    return h if pre_tanh else torch.tanh(h)


def image_location(ind_h: int, ind_w: int, steps_h: int, steps_w: int) -> str:
    """Location strings of the test sampler (utils.py:321-337)."""
    if steps_h == 1:
        s = "1st_row_last_row"
    elif ind_h == 0:
        s = "1st_row"
    elif ind_h == steps_h - 1:
        s = "last_row"
    else:
        s = "inter_row"
    if steps_w == 1:
        s += "_1st_col_last_col"
    elif ind_w == 0:
        s += "_1st_col"
    elif ind_w == steps_w - 1:
        s += "_last_col"
    else:
        s += "_inter_col"
    return s


def make_noise(cfg: GenCfg, total_h: int, total_w: int, seed: int, dtype=torch.float32):
    """Host-side noise for a total_h x total_w patch grid, drawn in the order of utils.py:306-308:
    z (1,z_dim,total_h*base+2,total_w*base+2) then, for SSM, level maps (1,1,total_h*r_i+4,total_w*r_i+4)."""
    g = torch.Generator().manual_seed(seed)
    b = cfg.base_res
    z = torch.randn(1, cfg.z_dim, total_h * b + 2, total_w * b + 2, generator=g).to(dtype)
    maps = None
    if cfg.type_norm == "SSM":
        maps = [torch.randn(1, cfg.map_dim, total_h * b * 2 ** i + 4, total_w * b * 2 ** i + 4, generator=g).to(dtype)
                for i in range(cfg.n_layers_G)]
    return z, maps


def sample_patch_by_patch(sd, cfg: GenCfg, out_h: int, out_w: int, z_full: Tensor,
                          maps_full: Optional[Sequence[Tensor]] = None, max_steps: Optional[int] = None) -> Tensor:
    """sample_from_gen_PatchByPatch_test (utils.py:258-397) on caller-supplied noise ("Oracle B").

    z_full / maps_full are the full-grid noise tensors the reference would draw at utils.py:228,246.
    max_steps bounds the number of sub-image steps (CPU-baseline sampling); the partial canvas
    is returned in that case."""
    geo = geometry(out_h, out_w, cfg)
    P, nph, npw, b = geo["P"], cfg.num_patches_h, cfg.num_patches_w, cfg.base_res
    sh, sw = geo["steps_h"], geo["steps_w"]
    z_sub = crop_windows(z_full, nph * b + 2, npw * b + 2, (npw - 1) * b)            # utils.py:232
    map_sub = None
    if cfg.type_norm == "SSM":
        map_sub = [crop_windows(m, nph * b * 2 ** i + 4, npw * b * 2 ** i + 4, (npw - 1) * b * 2 ** i)
                   for i, m in enumerate(maps_full)]                                # utils.py:246-253
    state = SeqState()
    rows, k = [], 0
    for ih in range(sh):
        row = []
        for iw in range(sw):
            loc = image_location(ih, iw, sh, sw)
            maps = None
            if map_sub is not None:                                                 # utils.py:345-351
                maps = [crop_windows(map_sub[i][[k]], b * 2 ** i + 4, b * 2 ** i + 4, b * 2 ** i)
                        for i in range(cfg.n_layers_G)]
            patches = forward_patches(sd, cfg, z_sub[[k]], maps, loc, state)
            sub = merge_patches(patches, nph, npw)
            keep_h = sub.shape[-2] if ih == sh - 1 else P * (nph - 1)                # utils.py:364-377
            keep_w = sub.shape[-1] if iw == sw - 1 else P * (npw - 1)
            row.append(sub[:, :, :keep_h, :keep_w])
            k += 1
            if max_steps is not None and k >= max_steps:
                return torch.cat(row, -1)
        rows.append(torch.cat(row, -1))
    full = torch.cat(rows, -2)
    return full[:, :, :out_h, :out_w]                                                # utils.py:395


# --------------------------------------------------------------------------------------------
# Formulation 2: one-shot on the merged grid (Oracle A; what the CUDA path computes)
# --------------------------------------------------------------------------------------------
def _conv_same(x: Tensor, sd, prefix: str, mode: str) -> Tensor:
    """3x3 conv of the merged image with a 1-px replicate/constant border (== LocalPadder + valid conv
    over all patches)."""
    return _conv(F.pad(x, (1, 1, 1, 1), mode), sd, prefix + "conv.")


def _ssm_merged(x: Tensor, m: Tensor, sd, prefix: str) -> Tensor:
    # valid convs on the over-sized merged map give exactly the merged modulation
    return ssm(x, m, sd, prefix)


def _attention_merged(x: Tensor, sd, patch: int) -> Tensor:
    """Per-patch attention on a merged (1,C,th*patch,tw*patch) tensor."""
    _, C, H, W = x.shape
    th, tw = H // patch, W // patch
    p = x.reshape(1, C, th, patch, tw, patch).permute(0, 2, 4, 1, 3, 5).reshape(th * tw, C, patch, patch)
    p = attention(p, sd)
    return p.reshape(1, th, tw, C, patch, patch).permute(0, 3, 1, 4, 2, 5).reshape(1, C, H, W)


def forward_merged(sd, cfg: GenCfg, z_full: Tensor, maps_full: Optional[Sequence[Tensor]] = None,
                   pre_tanh: bool = False, taps: Optional[dict] = None) -> Tensor:
    """One-shot Generator on the merged grid: z_full (1,z_dim,4*th+2,4*tw+2) -> (1,img_ch,th*P,tw*P).

    `taps`, if given, receives named intermediate activations (merged NCHW) for layer-wise tests."""
    mode = _pad_mode(cfg.outer_padding)
    maps = list(maps_full) if maps_full is not None else [None] * cfg.n_layers_G
    dt = z_full.dtype

    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    h = rec("start", _conv(z_full, sd, "start.conv."))                              # valid conv on the ringed z
    for k, (cin, cout) in enumerate(cfg.block_channels(), start=1):
        if k > 1:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
        m = None if maps[k - 1] is None else maps[k - 1].to(dt)
        pre = f"block{k}."
        a = _act(_norm(h, m, sd, pre + "bn1.", cfg), cfg.leak)
        t = rec(pre + "conv1", _conv_same(a, sd, pre + "conv1.", mode))
        a = _act(_norm(t, m, sd, pre + "bn2.", cfg), cfg.leak)
        t = _conv_same(a, sd, pre + "conv2.", mode)
        sc = h
        if cin != cout:
            if cfg.type_norm == "SSM":
                sc = ssm(sc, m, sd, pre + "bn3.")
            sc = _conv(sc, sd, pre + "conv3.")
        h = rec(pre + "out", t + sc)
        if k == 3 and cfg.attention:
            h = rec("attention", _attention_merged(h, sd, 4 * cfg.base_res))
    if cfg.type_norm == "BN":
        h = _bn_eval(h, sd, "bn.")
    h = _act(h, cfg.leak)
    h = rec("final_pre_tanh", _conv_same(h, sd, "final.", mode))
    return h if pre_tanh else torch.tanh(h)


# --------------------------------------------------------------------------------------------
# The non-local Generator (--padding_mode zeros): every conv is zero-padded, no patch structure
# --------------------------------------------------------------------------------------------
def _ssm_zeros(x: Tensor, m: Tensor, sd, prefix: str) -> Tensor:
    """StochasticSpatialModulation.forward with padding_mode='zeros' (p = 1 for mlp_shared and embed, models/layers.py:213-224):
    x (B,C,r,r), m (B,1,r,r)."""
    out = _bn_eval(x, sd, prefix + "bn.", affine=False)
    actv = F.relu(_conv(m.to(x.dtype), sd, prefix + "mlp_shared.0.", padding=1))
    emb = _conv(actv, sd, prefix + "embed.", padding=1)
    gamma, beta = emb.chunk(2, dim=1)
    return (1 + gamma) * out + beta


def forward_nonlocal(sd, cfg: GenCfg, z: Tensor, maps: Optional[Sequence[Tensor]] = None, pre_tanh: bool = False) -> Tensor:
    """ResidualPatchGenerator.forward (models/generators.py:86-124) built with padding_mode='zeros' (conv2d_lp -> conv3x3(..., p=1),
    models/layers.py:26-27): z (N,z_dim,h,w) -> (N,img_ch,h*2^(n-1),w*2^(n-1)); attention over the whole level-3 map."""
    maps = list(maps) if maps is not None else [None] * cfg.n_layers_G
    ssm_mode = cfg.type_norm == "SSM"

    def norm(x, m, prefix):
        return _ssm_zeros(x, m, sd, prefix) if ssm_mode else _bn_eval(x, sd, prefix)

    h = _conv(z, sd, "start.conv.", padding=1)
    for k, (cin, cout) in enumerate(cfg.block_channels(), start=1):
        if k > 1:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
        m = maps[k - 1]
        pre = f"block{k}."
        t = _conv(_act(norm(h, m, pre + "bn1."), cfg.leak), sd, pre + "conv1.conv.", padding=1)
        t = _conv(_act(norm(t, m, pre + "bn2."), cfg.leak), sd, pre + "conv2.conv.", padding=1)
        sc = h
        if cin != cout:
            if ssm_mode:
                sc = _ssm_zeros(sc, m, sd, pre + "bn3.")
            sc = _conv(sc, sd, pre + "conv3.")
        h = t + sc
        if k == 3 and cfg.attention:
            h = attention(h, sd)
    if not ssm_mode:
        h = _bn_eval(h, sd, "bn.")
    h = _conv(_act(h, cfg.leak), sd, "final.conv.", padding=1)
    return h if pre_tanh else torch.tanh(h)


def tile_process(z: Tensor, model, scale: int, tile_size: int = 32, tile_pad: int = 8) -> Tensor:
    """utils.tile_process (utils.py:401-470): overlapping latent tiles, centres pasted into the output."""
    n, _, height, width = z.shape
    out = None
    for y in range(math.ceil(height / tile_size)):
        for x in range(math.ceil(width / tile_size)):
            x0, y0 = x * tile_size, y * tile_size
            x1, y1 = min(x0 + tile_size, width), min(y0 + tile_size, height)
            x0p, x1p, y0p, y1p = max(x0 - tile_pad, 0), min(x1 + tile_pad, width), max(y0 - tile_pad, 0), min(y1 + tile_pad, height)
            t = model(z[:, :, y0p:y1p, x0p:x1p])
            if out is None:
                out = t.new_zeros((n, t.shape[1], height * scale, width * scale))
            ox, oy = (x0 - x0p) * scale, (y0 - y0p) * scale
            out[:, :, y0 * scale:y1 * scale, x0 * scale:x1 * scale] = t[:, :, oy:oy + (y1 - y0) * scale, ox:ox + (x1 - x0) * scale]
    return out


def image_to_patches(img: Tensor, P: int) -> Tensor:
    """(1,C,th*P,tw*P) -> (th*tw,C,P,P), row-major patch order (the Generator's return layout)."""
    return crop_windows(img, P, P, P)


# --------------------------------------------------------------------------------------------
# deterministic weights (state-dict layout of the reference)
# --------------------------------------------------------------------------------------------
def state_dict_spec(cfg: GenCfg) -> Dict[str, Tuple[int, ...]]:
    """Name -> shape of every tensor in the reference Generator's state_dict."""
    spec: Dict[str, Tuple[int, ...]] = {}

    def conv(name, co, ci, k):
        spec[name + "weight"] = (co, ci, k, k)
        spec[name + "bias"] = (co,)

    def bn(name, c, affine=True):
        if affine:
            spec[name + "weight"] = (c,)
            spec[name + "bias"] = (c,)
        spec[name + "running_mean"] = (c,)
        spec[name + "running_var"] = (c,)
        spec[name + "num_batches_tracked"] = ()

    def norm(name, c):
        if cfg.type_norm == "SSM":
            bn(name + "bn.", c, affine=False)
            conv(name + "mlp_shared.0.", 128, cfg.map_dim, 3)
            conv(name + "embed.", 2 * c, 128, 3)
        else:
            bn(name, c)

    conv("start.conv.", 8 * cfg.G_ch, cfg.z_dim, 3)
    for k, (ci, co) in enumerate(cfg.block_channels(), start=1):
        p = f"block{k}."
        conv(p + "conv1.conv.", co, ci, 3)
        conv(p + "conv2.conv.", co, co, 3)
        if ci != co:
            conv(p + "conv3.", co, ci, 1)
        norm(p + "bn1.", ci)
        norm(p + "bn2.", co)
        if ci != co and cfg.type_norm == "SSM":
            norm(p + "bn3.", ci)
    if cfg.type_norm == "BN":
        bn("bn.", cfg.final_ch)
    if cfg.attention:
        c = 2 * cfg.G_ch
        spec["attention.gamma"] = ()
        conv("attention.theta.", c // 8, c, 1)
        conv("attention.phi.", c // 8, c, 1)
        conv("attention.g.", c // 2, c, 1)
        conv("attention.o.", c, c // 2, 1)
    conv("final.conv.", cfg.img_ch, cfg.final_ch, 3)
    return spec


def make_state_dict(cfg: GenCfg, seed: int, stress: bool = False, orthogonal: bool = False) -> Dict[str, Tensor]:
    """Deterministic weights at the scale of the reference's init (utils.py:745-762): conv weights with
    the entry variance of an orthogonal matrix (1/max(rows, cols)), zero bias, BN weight ~ N(1, .02),
    running stats 0/1, attention.gamma = 0.  Gaussian by default so that a seed reproduces the same
    tensors on every machine (golden fixtures store only the seed); orthogonal=True runs the QR the
    reference's nn.init.orthogonal_ does.

    stress=True perturbs everything the stock init multiplies by 0 / identity:
    biases, BN running stats, attention.gamma = 0.5."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in state_dict_spec(cfg).items():
        leaf = name.rsplit(".", 1)[-1]
        is_bn_affine = leaf in ("weight", "bias") and len(shape) == 1 and (
            name.startswith("bn.") or (".bn" in name and ".conv" not in name and "mlp_shared" not in name
                                       and "embed" not in name))
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf == "gamma":
            t = torch.tensor(0.5 if stress else 0.0)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g) if stress else torch.ones(shape)
        elif is_bn_affine and leaf == "weight":
            t = 1.0 + 0.02 * torch.randn(shape, generator=g)
        elif is_bn_affine and leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif leaf == "weight":
            rows, cols = shape[0], shape[1] * shape[2] * shape[3]
            if orthogonal:
                flat = torch.randn(max(rows, cols), min(rows, cols), generator=g)
                q, r = torch.linalg.qr(flat)
                q = q * torch.sign(torch.diagonal(r)).unsqueeze(0)
                t = (q.t() if rows < cols else q).reshape(shape).contiguous()
            else:
                t = torch.randn(shape, generator=g) / math.sqrt(max(rows, cols))
        else:  # conv bias
            t = 0.05 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        sd[name] = t
    return sd
