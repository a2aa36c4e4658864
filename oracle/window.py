"""Window checks for grids the CPU oracle cannot run whole (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

The Generator is strictly local: an artificial border contaminates at most one patch, two with a trained
attention block (gamma != 0), so the one-shot oracle on a cropped z / map window with a two-patch margin reproduces the interior
of the window exactly.  Sides of the window that coincide with the border of the full grid need no margin (same outer padding).
Follows models/layers.py:145-173 (overlapping crops) and utils.py:221-256 (noise geometry).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import itg_oracle as O


def crop_noise(cfg: "O.GenCfg", z_full: torch.Tensor, maps_full: Optional[Sequence[torch.Tensor]], r0: int, r1: int, c0: int, c1: int):
    """Noise of the patch window rows [r0, r1) x cols [c0, c1), rings included (z: 1 px, maps: 2 px)."""
    b = cfg.base_res
    z = z_full[:, :, r0 * b:r1 * b + 2, c0 * b:c1 * b + 2].contiguous()
    maps = None
    if maps_full is not None:
        maps = [m[:, :, r0 * b * 2 ** i:r1 * b * 2 ** i + 4, c0 * b * 2 ** i:c1 * b * 2 ** i + 4].contiguous() for i, m in enumerate(maps_full)]
    return z, maps


def window_error(sd, cfg: "O.GenCfg", z_full: torch.Tensor, maps_full, img: torch.Tensor, th: int, tw: int, window: Tuple[int, int, int, int],
                 margin: int = 2, pre_tanh: bool = False) -> Tuple[float, Tuple[int, int, int, int]]:
    """max |img - oracle| over the trusted interior of the patch window (r0, r1, c0, c1) of a th x tw grid.
    img: (1, C, th*P, tw*P) image of the whole grid (any device).  Returns (error, trusted patch rectangle)."""
    r0, r1, c0, c1 = window
    z, maps = crop_noise(cfg, z_full, maps_full, r0, r1, c0, c1)
    with torch.no_grad():
        ref = O.forward_merged(sd, cfg, z, maps, pre_tanh=pre_tanh)
    P = cfg.patch_px
    t0 = r0 if r0 == 0 else r0 + margin
    t1 = r1 if r1 == th else r1 - margin
    l0 = c0 if c0 == 0 else c0 + margin
    l1 = c1 if c1 == tw else c1 - margin
    assert t1 > t0 and l1 > l0, "window too small for its margins"
    got = img[:, :, t0 * P:t1 * P, l0 * P:l1 * P].detach().float().cpu()
    want = ref[:, :, (t0 - r0) * P:(t1 - r0) * P, (l0 - c0) * P:(l1 - c0) * P]
    return (got - want).abs().max().item(), (t0, t1, l0, l1)


def standard_windows(th: int, tw: int, size: int = 6):
    """A corner, an edge, the interior and the opposite corner of a th x tw patch grid (clipped to the grid)."""
    sh, sw = min(size, th), min(size, tw)
    mid_r, mid_c = max(0, (th - sh) // 2), max(0, (tw - sw) // 2)
    wins = {"top-left corner": (0, sh, 0, sw), "right edge": (mid_r, mid_r + sh, tw - sw, tw), "interior": (mid_r, mid_r + sh, mid_c, mid_c + sw),
            "bottom-right corner": (th - sh, th, tw - sw, tw)}
    seen, out = set(), {}
    for k, w in wins.items():
        if w not in seen:
            seen.add(w)
            out[k] = w
    return out
