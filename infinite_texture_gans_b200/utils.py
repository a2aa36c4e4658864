"""Host-side sampler: mirror of the reference's utils.py surface for the patch-by-patch inference path.

`sample_from_gen_PatchByPatch_test` keeps the reference signature (utils.py:258-259) and noise draw
order (utils.py:228, 246: one `torch.randn` for the whole z grid, then one per SSM level), so the same
torch seed gives the same texture.  Two schedules:

* ``schedule='auto'`` (default): 'sequential' when the Generator has an attention block with gamma != 0 (a trained
  checkpoint: only then do the two schedules differ), 'oneshot' otherwise -- so the default call always reproduces
  the image the reference's sampler produces.
* ``schedule='oneshot'``: the whole total_h x total_w patch grid is one device-resident forward
  (what the training-time sampler utils.py:475-527 does with `LocalPadder.set_attributes(total_h, total_w)`;
  bit-for-bit what the sequential schedule produces unless attention.gamma != 0).  No patch
  is computed twice, nothing crosses the PCIe bus between layers.
* ``schedule='sequential'``: the shipped schedule -- num_patches_height x num_patches_width sub-images with
  stored halos (models/layers.py:103-143) and drop/regenerate of the last patch row / column
  (utils.py:364-377) -- for exact reproduction of trained checkpoints with attention.
"""
from __future__ import annotations

from math import ceil
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .layers import LocalPadder


# ------------------------------------------------------------------------------------------------
# geometry and patch <-> image plumbing (utils.py:294-303, 577-613, 658-742)
# ------------------------------------------------------------------------------------------------
def patch_grid_geometry(output_resolution_height: int, output_resolution_width: int, n_layers_G: int, base_res: int = 4,
                        num_patches_height: int = 3, num_patches_width: int = 3) -> Dict[str, int]:
    """steps / total patch counts of utils.py:294-303."""
    P = (2 ** (n_layers_G - 1)) * base_res
    if num_patches_height < 2 or num_patches_width < 2:
        raise ValueError("the sub-image needs at least 2 x 2 patches (one is dropped and regenerated per step)")
    steps_h = ceil((output_resolution_height / P - 1) / (num_patches_height - 1))
    steps_w = ceil((output_resolution_width / P - 1) / (num_patches_width - 1))
    if steps_h < 1 or steps_w < 1:
        raise ValueError(f"output {output_resolution_height}x{output_resolution_width} must exceed one patch ({P} px) per "
                         "side (utils.py:317-395 fails with an UnboundLocalError there)")
    return dict(P=P, steps_h=steps_h, steps_w=steps_w, total_h=steps_h * (num_patches_height - 1) + 1,
                total_w=steps_w * (num_patches_width - 1) + 1)


def image_location_of(ind_h: int, ind_w: int, steps_h: int, steps_w: int) -> str:
    """Location strings of utils.py:321-337."""
    row = "1st_row_last_row" if steps_h == 1 else ("1st_row" if ind_h == 0 else ("last_row" if ind_h == steps_h - 1 else "inter_row"))
    col = "_1st_col_last_col" if steps_w == 1 else ("_1st_col" if ind_w == 0 else ("_last_col" if ind_w == steps_w - 1 else "_inter_col"))
    return row + col


def crop_images(images: torch.Tensor, crop_height: int, crop_width: int, stride: int, device=None) -> torch.Tensor:
    """Sliding-window crops, row-major, images outermost (utils.py:658-742): (N,C,H,W) -> (N*P,C,ch,cw) fp32."""
    N, C, H, W = images.shape
    win = images.unfold(2, crop_height, stride).unfold(3, crop_width, stride)     # (N,C,ny,nx,ch,cw)
    out = win.permute(0, 2, 3, 1, 4, 5).reshape(-1, C, crop_height, crop_width).float()
    return out if device is None else out.to(device)


def merge_patches_into_image(patches: torch.Tensor, num_patches_height: int, num_patches_width: int, device=None) -> torch.Tensor:
    """Inverse of a non-overlapping crop (utils.py:577-613): (B,C,h,w) -> (B/(nph*npw), C, nph*h, npw*w) fp32."""
    B, C, h, w = patches.shape
    n = B // (num_patches_height * num_patches_width)
    x = patches.reshape(n, num_patches_height, num_patches_width, C, h, w).permute(0, 3, 1, 4, 2, 5)
    out = x.reshape(n, C, num_patches_height * h, num_patches_width * w).float()
    return out if device is None else out.to(device)


def draw_noise(num_images: int, z_dim: int, base_res: int, n_layers_G: int, map_dim: int, type_norm: str, total_h: int,
               total_w: int) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]]]:
    """Full-grid noise in the reference's draw order on the host RNG (utils.py:228 then :246 per level)."""
    z = torch.randn(num_images, z_dim, total_h * base_res + 2, total_w * base_res + 2)
    maps = None
    if type_norm == "SSM":
        maps = [torch.randn(num_images, map_dim, total_h * base_res * 2 ** i + 4, total_w * base_res * 2 ** i + 4)
                for i in range(n_layers_G)]
    return z, maps


def noise_window(shape_full: Tuple[int, int, int], window: Tuple[int, int, int, int], seed: int, field: int, device) -> torch.Tensor:
    """Window (y0, y1, x0, x1) of the counter-based standard-normal field `field` of shape (C, Hf, Wf) (include/itg.h itg_noise_normal),
    generated on `device`.  A pure function of (seed, field, position): any two windows agree where they overlap."""
    C, Hf, Wf = shape_full
    y0, y1, x0, x1 = window
    out = torch.empty((C, y1 - y0, x1 - x0), dtype=torch.float32, device=device)
    lib = L.load()
    with torch.cuda.device(out.device):
        L.check(lib.itg_noise_normal(L.ptr(out), C, y1 - y0, x1 - x0, y0, x0, Hf, Wf, int(seed) & 0xFFFFFFFFFFFFFFFF, field, L.stream_ptr()))
    return out


def draw_noise_device(cfg, total_h: int, total_w: int, seed: int, rows: Optional[Tuple[int, int]] = None, device="cuda"):
    """Device-side counterpart of `draw_noise` / utils.build_z + build_maps (utils.py:221-256) for large grids: the noise of patch rows
    `rows` = (r0, r1) (default: all) of a total_h x total_w grid, generated on the device from the counter-based generator -- no host
    draw, no host -> device copy, and every rank / sub-image sees the same field.  Same distribution as the reference's torch.randn
    draws, different stream.  Returns (z (z_dim, rows*b+2, W), [map_i (rows*r_i+4, W_i)] or None), the layout
    `Plan.set_inputs` / `RowBandSampler.set_band_noise` take."""
    b = cfg.base_res
    r0, r1 = rows if rows is not None else (0, total_h)
    z = noise_window((cfg.z_dim, total_h * b + 2, total_w * b + 2), (r0 * b, r1 * b + 2, 0, total_w * b + 2), seed, 0, device)
    maps = None
    if cfg.type_norm == "SSM":
        maps = []
        for i in range(cfg.n_layers_G):
            r = b * 2 ** i
            maps.append(noise_window((1, total_h * r + 4, total_w * r + 4), (r0 * r, r1 * r + 4, 0, total_w * r + 4), seed, 1 + i, device)[0])
    return z, maps


def build_z(num_images=1, z_dim=128, base_res=4, num_patches_height=3, num_patches_width=3, total_num_patches_height=3,
            total_num_patches_width=3, device="cpu"):
    """utils.py:221-234: overlapping sub-image crops of the full latent grid (with its random 1-px ring)."""
    z, _ = draw_noise(num_images, z_dim, base_res, 0, 1, "BN", total_num_patches_height, total_num_patches_width)
    return crop_images(z.to(device), num_patches_height * base_res + 2, num_patches_width * base_res + 2,
                       (num_patches_width - 1) * base_res)


def build_maps(num_images=1, map_dim=1, n_layers_G=4, base_res=4, num_patches_height=3, num_patches_width=3,
               total_num_patches_height=3, total_num_patches_width=3, device="cpu"):
    """utils.py:237-256: per level, overlapping sub-image crops of the full noise map (4-px over-size)."""
    out = []
    for i in range(n_layers_G):
        r = base_res * 2 ** i
        m = torch.randn(num_images, map_dim, total_num_patches_height * r + 4, total_num_patches_width * r + 4).to(device)
        out.append(crop_images(m, num_patches_height * r + 4, num_patches_width * r + 4, (num_patches_width - 1) * r))
    return out


def _unwrap(netG):
    return netG.module if isinstance(netG, nn.DataParallel) else netG


# ------------------------------------------------------------------------------------------------
# samplers
# ------------------------------------------------------------------------------------------------
def generate_full_grid(netG, z_full: torch.Tensor, maps_full: Optional[Sequence[torch.Tensor]] = None, graph: bool = False,
                       pre_tanh: bool = False) -> torch.Tensor:
    """One-shot forward of a whole patch grid.  z_full: (1, z_dim, th*b+2, tw*b+2) host or device fp32;
    maps_full: per level (1, 1, th*r+4, tw*r+4).  Returns the device-resident (1, img_ch, th*P, tw*P) image
    (the engine's output buffer: valid until the next call on the same grid size).  pre_tanh=True returns the final conv's
    fp32 output before the tanh instead (parity checks: a tenth of the outputs of a trained Generator saturate)."""
    G = _unwrap(netG)
    b = G.cfg.base_res
    th, tw = (z_full.shape[-2] - 2) // b, (z_full.shape[-1] - 2) // b
    if z_full.shape[0] != 1:
        raise ValueError("one texture per call (utils.py:341 ignores num_images as well)")
    eng = G.engine()
    maps = None if maps_full is None else [m[0, 0] for m in maps_full]
    if pre_tanh:
        return eng.forward_pre_tanh(z_full[0], maps, th=th, tw=tw)
    return eng.forward(z_full[0], maps, th=th, tw=tw, img_layout=L.IMG_MERGED, graph=graph)


def resolve_schedule(netG, schedule: str = "auto") -> str:
    """'auto' -> the schedule that reproduces the reference's sampler: the one-shot pass equals the shipped sequential
    3x3 schedule bit for bit unless the attention block contributes (attention.gamma != 0), in which case
    the outer patches the reference drops and regenerates differ and only 'sequential' matches."""
    if schedule not in ("auto", "oneshot", "sequential"):
        raise ValueError("schedule must be 'auto', 'oneshot' or 'sequential'")
    if schedule != "auto":
        return schedule
    att = getattr(_unwrap(netG), "attention", False)
    gamma = getattr(att, "gamma", None) if att is not False else None
    return "sequential" if gamma is not None and float(gamma.detach().float().cpu()) != 0.0 else "oneshot"


def sample_from_gen_PatchByPatch_test(netG, z_dim=128, base_res=4, map_dim=1, num_images=1, num_patches_height=3,
                                      num_patches_width=3, device="cpu", output_resolution_height=384,
                                      output_resolution_width=384, schedule: str = "auto", noise=None,
                                      return_on_device: bool = False, graph: bool = False) -> torch.Tensor:
    """Generate one (1, img_ch, H, W) texture (utils.py:258-397).  Returns a host fp32 tensor like the
    reference (each of its sub-images is `.cpu()`-ed, utils.py:360) unless return_on_device.

    schedule: 'auto' (default; see resolve_schedule), 'oneshot' or 'sequential'.
    noise: optional (z_full, maps_full) to use instead of drawing from the global host RNG.
    graph: replay the one-shot launch list from a CUDA graph (captured on first use per grid size).
    base_res must be the Generator's own (the reference's sampler trusts its argument, utils.py:258; its Generator is fully
    convolutional and ignores base_res, here the launch plan is built for the Generator's value)."""
    G = _unwrap(netG)
    n_layers_G, type_norm = G.n_layers_G, G.type_norm
    if base_res != G.cfg.base_res:
        raise ValueError(f"base_res={base_res} does not match the Generator's base_res={G.cfg.base_res}: pass the checkpoint's "
                         "args.base_res (test_sample.py:68 relies on the default 4)")
    schedule = resolve_schedule(netG, schedule)
    geo = patch_grid_geometry(output_resolution_height, output_resolution_width, n_layers_G, base_res,
                              num_patches_height, num_patches_width)
    P, th, tw = geo["P"], geo["total_h"], geo["total_w"]
    if noise is None:
        z_full, maps_full = draw_noise(num_images, z_dim, base_res, n_layers_G, map_dim, type_norm, th, tw)
    else:
        z_full, maps_full = noise
    z_full = z_full[:1]                      # utils.py:341 only ever uses the first image's crops
    if maps_full is not None:
        maps_full = [m[:1] for m in maps_full]
    H, W = output_resolution_height, output_resolution_width
    if schedule == "oneshot":
        img = generate_full_grid(netG, z_full, maps_full, graph=graph)[:, :, :H, :W]
        return img if return_on_device else img.cpu()

    # ---- the shipped schedule (utils.py:317-392) with on-device halos and on-device assembly ----
    dev = next(G.parameters()).device
    nph, npw, b = num_patches_height, num_patches_width, base_res
    saved = (LocalPadder.num_patches_h, LocalPadder.num_patches_w)
    LocalPadder.num_patches_h, LocalPadder.num_patches_w = nph, npw
    try:
        canvas = torch.empty((1, G.img_ch, th * P, tw * P), dtype=torch.float32, device=dev)
        z_dev = z_full.to(dev)
        maps_dev = None if maps_full is None else [m.to(dev) for m in maps_full]
        sh, sw = geo["steps_h"], geo["steps_w"]
        for ih in range(sh):
            for iw in range(sw):
                loc = image_location_of(ih, iw, sh, sw)
                py, px = ih * (nph - 1), iw * (npw - 1)                   # top-left patch of this sub-image
                z_sub = z_dev[:, :, py * b:(py + nph) * b + 2, px * b:(px + npw) * b + 2]
                maps = None
                if maps_dev is not None:
                    maps = [m[:, :, py * b * 2 ** i:(py + nph) * b * 2 ** i + 4, px * b * 2 ** i:(px + npw) * b * 2 ** i + 4]
                            for i, m in enumerate(maps_dev)]
                patches = G(z_sub.contiguous(), maps, loc)
                sub = merge_patches_into_image(patches, nph, npw)
                kh = nph * P if ih == sh - 1 else (nph - 1) * P          # utils.py:364-377
                kw = npw * P if iw == sw - 1 else (npw - 1) * P
                canvas[:, :, py * P:py * P + kh, px * P:px * P + kw] = sub[:, :, :kh, :kw]
    finally:
        LocalPadder.num_patches_h, LocalPadder.num_patches_w = saved
    img = canvas[:, :, :H, :W]
    return img if return_on_device else img.cpu()


VIRTUAL_ROWS = 1 << 24      # patch rows of the counter-based noise field a seeded stream draws from (2^31 pixels of height at P = 128)


def stream_texture_rows(netG, output_resolution_width: int, *, seed: Optional[int] = None, noise_rows=None, steps: Optional[int] = None,
                        num_patches_height: int = 3, num_patches_width: int = 3):
    """Top-to-bottom generator of a texture of UNBOUNDED height: the shipped sub-image schedule (utils.py:317-392) run one
    sub-image ROW per iteration, yielding the finished patch rows of each as a device tensor (1, img_ch, rows * P, W) -- (nph - 1) patch rows per
    iteration, nph for the last one.  Device memory is O(one sub-image row) whatever the height: the noise of nph patch rows, one band canvas
    and LocalPadder's halo state (the row above + the sub-image to the left, layers.py:103-143); nothing of the texture is kept.

    steps: number of sub-image rows; the last one is generated with the 'last_row' locations (bottom outer padding).  None = endless
           (every row after the first is an 'inter_row'; stop iterating when enough has been produced).
    Noise: `seed` draws patch rows on demand from the device-side counter-based field (draw_noise_device on a virtual grid of
           VIRTUAL_ROWS patch rows: any two streams with the same seed and width agree row for row), or `noise_rows(ih)` returns
           (z, maps) of patch rows [ih * (nph - 1), ih * (nph - 1) + nph) in draw_noise's 4-D layout (host or device).
    With `noise_rows` slicing a full grid and `steps` = its steps_h, the concatenated output equals
    sample_from_gen_PatchByPatch_test(..., schedule='sequential') bit for bit (same launches in the same order)."""
    G = _unwrap(netG)
    cfg = G.cfg
    b, nph, npw = cfg.base_res, num_patches_height, num_patches_width
    geo = patch_grid_geometry(2 * cfg.patch_px, output_resolution_width, cfg.n_layers_G, b, nph, npw)
    P, tw, sw = geo["P"], geo["total_w"], geo["steps_w"]
    if (seed is None) == (noise_rows is None):
        raise ValueError("pass either seed= (device-side counter-based noise) or noise_rows= (a callable returning each row's noise)")
    if steps is not None and steps < 1:
        raise ValueError("steps must be >= 1")
    dev = next(G.parameters()).device

    def rows_of(ih):
        if noise_rows is not None:
            z, maps = noise_rows(ih)
            return z[:1].to(dev), None if maps is None else [m[:1].to(dev) for m in maps]
        py = ih * (nph - 1)
        if py + nph > VIRTUAL_ROWS:
            raise RuntimeError("the seeded noise field is exhausted (2^24 patch rows)")
        z, maps = draw_noise_device(cfg, VIRTUAL_ROWS, tw, seed, rows=(py, py + nph), device=dev)
        return z.unsqueeze(0), None if maps is None else [m[None, None] for m in maps]

    saved = (LocalPadder.num_patches_h, LocalPadder.num_patches_w)
    LocalPadder.num_patches_h, LocalPadder.num_patches_w = nph, npw
    try:
        ih = 0
        while steps is None or ih < steps:
            last = steps is not None and ih == steps - 1
            z_rows, maps_rows = rows_of(ih)
            keep = nph if last else nph - 1
            band = torch.empty((1, G.img_ch, keep * P, tw * P), dtype=torch.float32, device=dev)
            row = "1st_row_last_row" if (last and ih == 0) else ("1st_row" if ih == 0 else ("last_row" if last else "inter_row"))
            for iw in range(sw):
                col = "_1st_col_last_col" if sw == 1 else ("_1st_col" if iw == 0 else ("_last_col" if iw == sw - 1 else "_inter_col"))
                px = iw * (npw - 1)
                z_sub = z_rows[:, :, :, px * b:(px + npw) * b + 2]
                maps = None
                if maps_rows is not None:
                    maps = [m[:, :, :, px * b * 2 ** i:(px + npw) * b * 2 ** i + 4].contiguous() for i, m in enumerate(maps_rows)]
                sub = merge_patches_into_image(G(z_sub.contiguous(), maps, row + col), nph, npw)
                kw = npw * P if iw == sw - 1 else (npw - 1) * P              # utils.py:364-377
                band[:, :, :, px * P:px * P + kw] = sub[:, :, :keep * P, :kw]
            yield band[:, :, :, :output_resolution_width]
            ih += 1
    finally:
        LocalPadder.num_patches_h, LocalPadder.num_patches_w = saved


def tile_process(img: torch.Tensor, model, scale: int = 4, tile_size: int = 32, tile_pad: int = 8) -> torch.Tensor:
    """utils.tile_process (utils.py:401-470): run `model` on overlapping tiles of the latent `img` (N, C, h, w) and paste the centres
    into the (N, 3, h*scale, w*scale) output.  Same tiling arithmetic; the output stays on the device of the tiles."""
    from math import ceil as _ceil
    batch, _, height, width = img.shape
    out = None
    tiles_x, tiles_y = _ceil(width / tile_size), _ceil(height / tile_size)
    for y in range(tiles_y):
        for x in range(tiles_x):
            x0, y0 = x * tile_size, y * tile_size
            x1, y1 = min(x0 + tile_size, width), min(y0 + tile_size, height)
            x0p, x1p = max(x0 - tile_pad, 0), min(x1 + tile_pad, width)
            y0p, y1p = max(y0 - tile_pad, 0), min(y1 + tile_pad, height)
            tile = model(img[:, :, y0p:y1p, x0p:x1p])
            if out is None:
                out = tile.new_zeros((batch, tile.shape[1], height * scale, width * scale))
            ox, oy = (x0 - x0p) * scale, (y0 - y0p) * scale
            out[:, :, y0 * scale:y1 * scale, x0 * scale:x1 * scale] = tile[:, :, oy:oy + (y1 - y0) * scale, ox:ox + (x1 - x0) * scale]
    return out


def sample_from_gen(netG, z_dim=128, base_res=4, map_dim=1, num_images=1, tiles=False, device="cpu", noise=None):
    """utils.sample_from_gen (utils.py:530-575): the sampler of the NON-local Generator (--padding_mode zeros): z ~ N(0, 1) of
    (num_images, z_dim, base_res, base_res), per-level SSM maps of (num_images, 1, r, r), one forward -- or, with tiles=True,
    `tile_process(z, netG, 2^(n-1), 32, 16)` (test_sample.py:70-73).  Same draw order on the host RNG as the reference.
    noise: optional (z, maps) to use instead of drawing."""
    G = _unwrap(netG)
    n_layers_G, type_norm = G.n_layers_G, G.type_norm
    if noise is None:
        z = torch.randn(num_images, z_dim, base_res, base_res)
        maps = None
        if type_norm == "SSM":
            maps = [torch.randn(num_images, map_dim, base_res * 2 ** i, base_res * 2 ** i) for i in range(n_layers_G)]
    else:
        z, maps = noise
    if maps is None:
        maps = [None] * n_layers_G
    if tiles:
        if type_norm == "SSM":
            raise ValueError("tile_process feeds the model the latent only (utils.py:443): SSM Generators cannot be tiled")
        return tile_process(z, netG, 2 ** (n_layers_G - 1), 32, 16)
    return netG(z, maps, image_location="1st_row_1st_col")


# ------------------------------------------------------------------------------------------------
# output stage
# ------------------------------------------------------------------------------------------------
def image_to_uint8(img: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device-side output stage of test_sample.py:75-79: the bytes `torchvision.utils.save_image(img * 0.5 + 0.5, ...)` would write,
    i.e. `(img * 0.5 + 0.5).mul(255).add_(0.5).clamp_(0, 255).to(uint8)` as an interleaved (H, W, C) uint8 device tensor, bit-identical
    to the host computation.  img: (1, C, H, W) or (C, H, W) fp32 CUDA tensor; row / channel strides are honoured (cropped views of the
    Generator's output buffer), the innermost stride must be 1."""
    x = img[0] if img.dim() == 4 else img
    if x.dim() != 3 or x.dtype != torch.float32 or not x.is_cuda or x.stride(2) != 1:
        raise ValueError("image_to_uint8 takes a (1, C, H, W) / (C, H, W) fp32 CUDA tensor with unit innermost stride")
    c, h, w = x.shape
    if out is None:
        out = torch.empty((h, w, c), dtype=torch.uint8, device=x.device)
    elif tuple(out.shape) != (h, w, c) or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != x.device:
        raise ValueError(f"out must be a contiguous ({h}, {w}, {c}) uint8 tensor on {x.device}")
    lib = L.load()
    L.check(lib.itg_image_to_u8(x.data_ptr(), c, h, w, x.stride(1), x.stride(0), out.data_ptr(), L.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------
# streaming: many textures, copies overlapped with compute
# ------------------------------------------------------------------------------------------------
class HostOutputPipe:
    """Multi-buffered device -> pinned-host hand-off of finished images.  `push(img)` snapshots the device image
    (device-to-device, on the current stream) and starts its copy to pinned host memory on a side stream, so the next
    Generator pass runs while the previous image crosses PCIe; `wait(slot)` blocks until that image is on the host.
    A slot's host tensor is overwritten `depth` pushes later."""

    def __init__(self, shape, device, depth: int = 3, dtype: torch.dtype = torch.float32):
        self.depth, self.dtype = depth, dtype
        self.copy_stream = torch.cuda.Stream(device)
        self.stage = [torch.empty(shape, dtype=dtype, device=device) for _ in range(depth)]
        self.host = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]      # image of the slot is complete in self.stage
        self.done = [torch.cuda.Event() for _ in range(depth)]       # ... and in self.host
        self.k = 0

    def push(self, img: torch.Tensor) -> int:
        s = self.k % self.depth
        self.k += 1
        cur = torch.cuda.current_stream()
        if self.k > self.depth:
            cur.wait_event(self.done[s])                             # the slot's previous image has left the device
        if self.dtype == torch.uint8:
            image_to_uint8(img, out=self.stage[s])                   # quantise on the device: 1 byte per sample crosses PCIe
        else:
            self.stage[s].copy_(img)                                 # after this the engine may overwrite its output buffer
        self.ready[s].record(cur)
        self.copy_stream.wait_event(self.ready[s])
        with torch.cuda.stream(self.copy_stream):
            self.host[s].copy_(self.stage[s], non_blocking=True)
            self.done[s].record(self.copy_stream)
        return s

    def wait(self, slot: int) -> torch.Tensor:
        self.done[slot].synchronize()
        return self.host[slot]


class NoiseUploader:
    """Host -> device upload of the NEXT texture's noise on a side stream while the current pass computes: two device
    staging slots per plan; `upload(noise)` starts the copies, `feed()` makes the current stream wait for the oldest
    pending upload and moves it (device to device) into the plan's static input buffers."""

    def __init__(self, plan, device):
        self.plan = plan
        self.stream = torch.cuda.Stream(device)
        self.z = [torch.empty_like(plan.z_in) for _ in range(2)]
        self.maps = [[torch.empty_like(m) for m in plan.maps_in] for _ in range(2)]
        self.uploaded = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [None, None]
        self.head = self.tail = 0                        # uploads started / fed

    def upload(self, noise) -> None:
        z_full, maps_full = noise
        s = self.head % 2
        if self.head - self.tail >= 2:
            raise RuntimeError("NoiseUploader: both staging slots are pending")
        self.head += 1
        if len(self.maps[s]) and (maps_full is None or len(maps_full) < len(self.maps[s])):
            raise ValueError("SSM Generator needs one noise map per level (utils.py:237-256)")
        if self.consumed[s] is not None:
            self.stream.wait_event(self.consumed[s])     # the slot's previous content has been moved into the plan
        with torch.cuda.stream(self.stream):
            z = z_full[0] if z_full.dim() == 4 else z_full
            if tuple(z.shape) != tuple(self.z[s].shape):
                raise ValueError(f"z has shape {tuple(z.shape)}, the {self.plan.th}x{self.plan.tw} patch grid needs {tuple(self.z[s].shape)}")
            self.z[s].copy_(z, non_blocking=True)
            for dst, m in zip(self.maps[s], maps_full or ()):
                dst.copy_(m[0, 0] if m.dim() == 4 else m, non_blocking=True)
            self.uploaded[s].record(self.stream)

    def feed(self) -> None:
        s = self.tail % 2
        self.tail += 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self.uploaded[s])
        self.plan.set_inputs(self.z[s], self.maps[s] if self.maps[s] else None)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[s] = ev


def generate_textures(netG, noises, output_resolution_height: int, output_resolution_width: int, base_res: int = 4,
                      num_patches_height: int = 3, num_patches_width: int = 3, graph: bool = True, out_format: str = "float32"):
    """Iterator over host-resident (1, img_ch, H, W) fp32 textures, one per element of `noises`
    (each `(z_full, maps_full)` as drawn by `draw_noise`, ideally in pinned memory).  Same result per texture as
    `sample_from_gen_PatchByPatch_test(..., noise=...)` with the one-shot schedule; the difference is that texture k's
    copy to the host and texture k+2's noise upload overlap the Generator pass of texture k+1.  A yielded tensor is a view of a pinned staging buffer
    and stays valid until the iterator is advanced again.

    out_format='uint8' yields (H, W, img_ch) uint8 images instead: the bytes test_sample.py's `save_image(img * 0.5 + 0.5, ...)`
    writes (`image_to_uint8`), quantised on the device so that a quarter of the bytes cross PCIe."""
    if out_format not in ("float32", "uint8"):
        raise ValueError("out_format must be 'float32' or 'uint8'")
    G = _unwrap(netG)
    geo = patch_grid_geometry(output_resolution_height, output_resolution_width, G.n_layers_G, base_res,
                              num_patches_height, num_patches_width)
    H, W = output_resolution_height, output_resolution_width
    dev = next(G.parameters()).device
    pipes = G.__dict__.setdefault("_host_pipes", {})     # pinned staging buffers are expensive to allocate: keep them per output size
    pipe = pipes.get((H, W, dev, out_format))
    if pipe is None:
        pipe = pipes[(H, W, dev, out_format)] = (HostOutputPipe((H, W, G.img_ch), dev, dtype=torch.uint8) if out_format == "uint8"
                                                 else HostOutputPipe((1, G.img_ch, H, W), dev))
    it = iter(noises)
    nxt = next(it, None)
    if nxt is None:
        return
    b = G.cfg.base_res
    th, tw = (nxt[0].shape[-2] - 2) // b, (nxt[0].shape[-1] - 2) // b
    eng = G.engine()
    plan = eng.plan(th, tw, L.IMG_MERGED)
    ups = G.__dict__.setdefault("_noise_uploaders", {})
    up = ups.get((th, tw, dev))
    if up is None or up.plan is not plan:
        up = ups[(th, tw, dev)] = NoiseUploader(plan, dev)
    up.head = up.tail = 0
    with torch.cuda.device(dev):                         # (never across a yield: the consumer keeps its own current device)
        up.upload(nxt)
    in_flight = []                                       # at most depth - 1 images between the Generator pass and the consumer
    while nxt is not None:
        with torch.cuda.device(dev):
            up.feed()                                    # this texture's noise: staged on the device during the previous pass
            img = eng.replay(th, tw, L.IMG_MERGED) if graph else plan.run()
            nxt = next(it, None)
            if nxt is not None:
                up.upload(nxt)                           # the next texture's noise crosses PCIe while this pass computes
            in_flight.append(pipe.push(img[:, :, :H, :W]))
        if len(in_flight) == pipe.depth:
            yield pipe.wait(in_flight.pop(0))
    for slot in in_flight:
        yield pipe.wait(slot)
