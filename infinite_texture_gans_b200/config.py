"""Constructor arguments of the Generator on this path (reference models/generators.py:25-28) and the
derived per-level geometry.  Pure host-side bookkeeping; no torch dependency."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class GenConfig:
    z_dim: int = 128
    G_ch: int = 64
    base_res: int = 4
    n_layers_G: int = 4
    attention: bool = True
    img_ch: int = 3
    leak: float = 0.0
    type_norm: str = "BN"          # 'BN' | 'SSM'   (--type_norm_G, utils.py:69)
    map_dim: int = 1
    outer_padding: str = "replicate"   # 'replicate' | 'constant' ('zeros' is accepted as an alias)
    padding_mode: str = "local"        # 'local' (patch grid, LocalPadder) | 'zeros' (the non-local Generator: whole image, zero-padded convs)

    def __post_init__(self):
        if self.n_layers_G not in (4, 5, 6):
            raise ValueError(f"n_layers_G must be 4, 5 or 6 (models/generators.py:61-74), got {self.n_layers_G}")
        if self.type_norm not in ("BN", "SSM"):
            raise ValueError(f"type_norm must be 'BN' or 'SSM', got {self.type_norm!r}")
        if self.outer_padding not in ("replicate", "constant", "zeros"):
            raise ValueError(f"outer_padding must be 'replicate' or 'constant', got {self.outer_padding!r}")
        if self.padding_mode not in ("local", "zeros"):
            raise ValueError(f"padding_mode must be 'local' or 'zeros' (models/layers.py:19-27), got {self.padding_mode!r}")
        if self.map_dim != 1:
            raise ValueError("map_dim != 1 is not supported (test_sample.py:56 hard-codes map_dim=1)")
        if self.img_ch < 1 or self.img_ch > 8:
            raise ValueError("img_ch must be in 1..8")

    @property
    def patch_px(self) -> int:
        """Pixels per patch side at full resolution (utils.py:294)."""
        return self.base_res * 2 ** (self.n_layers_G - 1)

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

    def level_res(self, k: int) -> int:
        """Per-patch resolution at which block k (1-based) computes."""
        return self.base_res * 2 ** (k - 1)

    @property
    def border_is_replicate(self) -> bool:
        return self.outer_padding == "replicate" and self.padding_mode == "local"     # zero-padded convs: the frame is zeros

    @property
    def nonlocal_mode(self) -> bool:
        return self.padding_mode == "zeros"


def flops_per_patch(cfg: GenConfig) -> float:
    """Algorithmic FLOPs (2*MAC) of the reference's dense arithmetic for ONE patch: 3x3 convs,
    1x1 shortcuts, attention 1x1s + scores + PV, SSM convs; elementwise work excluded."""
    mac = 0.0
    r = cfg.base_res
    mac += 9 * cfg.z_dim * 8 * cfg.G_ch * r * r
    for k, (ci, co) in enumerate(cfg.block_channels(), start=1):
        r = cfg.level_res(k)
        px = r * r
        mac += 9 * ci * co * px + 9 * co * co * px
        if ci != co:
            mac += ci * co * px
        if cfg.type_norm == "SSM":
            norms = [ci, co] + ([ci] if ci != co else [])
            for c in norms:
                mac += 9 * 1 * 128 * (r + 2) * (r + 2) + 9 * 128 * 2 * c * px
        if k == 3 and cfg.attention:
            c = 2 * cfg.G_ch
            mac += px * c * (c // 8) * 2 + px * c * (c // 2) + px * (c // 2) * c     # theta, phi, g, o
            mac += px * (px // 4) * (c // 8) + px * (px // 4) * (c // 2)             # scores, PV
    mac += 9 * cfg.final_ch * cfg.img_ch * cfg.patch_px ** 2
    return 2.0 * mac
