"""Drop-in `ResidualPatchGenerator` for the patch-by-patch inference path.

Same constructor, same `forward(z, maps=None, image_location='1st_row_1st_col')`, same state-dict keys and
same return layout as the reference class (models/generators.py:4-124), so `test_sample.py`-style callers
(`load_state_dict`, `.eval()`, `.to(device)`, `utils.sample_from_gen_PatchByPatch_test(netG, ...)`) work
unchanged.  The module tree below only *holds* the parameters; the forward packs them once into the
operand layouts of libitg_b200.so (packing.py) and runs the fused launch plan of engine.py on the GPU.
There is no PyTorch or CPU fallback: a forward without the CUDA library raises.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from .config import GenConfig
from .engine import Engine
from .halo import SequentialHalo
from .layers import Attention, LocalPadder, ResBlockGenerator, conv2d_lp, init_weight


class ResidualPatchGenerator(nn.Module):
    """See models/generators.py:6-23 for the argument meanings.

    Extra keyword (defaults keep the reference behaviour):
      precision : 'fp16' (tcgen05 16-bit mode, default: image within 2e-2 of the fp32 reference) or 'fp32' (exact mode, within 1e-3: fp32 tensors, split-precision tensor-core convs).
    """

    def __init__(self, z_dim=128, G_ch=64, base_res=4, n_layers_G=4, attention=True, img_ch=3, leak=0, SN=False,
                 type_norm="BN", map_dim=1, padding_mode="local", outer_padding="replicate", num_patches_h=3,
                 num_patches_w=3, padding_size=1, conv_reduction=2, precision="fp16"):
        super().__init__()
        if padding_mode not in ("local", "zeros"):
            raise ValueError("padding_mode must be 'local' or 'zeros' (models/layers.py:19-27)")
        if SN:
            raise NotImplementedError("spectral-norm checkpoints are not loadable by test_sample.py either (SN=False, :56)")
        self.z_dim, self.base_ch, self.base_res, self.n_layers_G = z_dim, G_ch, base_res, n_layers_G
        self.img_ch, self.leak, self.SN, self.type_norm, self.map_dim = img_ch, leak, SN, type_norm, map_dim
        self.padding_mode, self.outer_padding = padding_mode, outer_padding
        self.num_patches_h, self.num_patches_w = num_patches_h, num_patches_w
        self.padding_size, self.conv_reduction = padding_size, conv_reduction
        self.precision = precision
        self.cfg = GenConfig(z_dim=z_dim, G_ch=G_ch, base_res=base_res, n_layers_G=n_layers_G, attention=bool(attention),
                             img_ch=img_ch, leak=float(leak), type_norm=type_norm, map_dim=map_dim,
                             outer_padding="constant" if outer_padding == "zeros" else outer_padding, padding_mode=padding_mode)
        # class-level, like the reference (models/generators.py:49-50)
        LocalPadder.set_attributes(num_patches_h=num_patches_h, num_patches_w=num_patches_w, outer_padding=outer_padding,
                                   padding_size=padding_size, conv_reduction=conv_reduction)

        self.start = conv2d_lp(z_dim, G_ch * 8)
        for k, (ci, co) in enumerate(self.cfg.block_channels(), start=1):
            setattr(self, f"block{k}", ResBlockGenerator(ci, co, type_norm, map_dim))
        if type_norm == "BN":
            self.bn = nn.BatchNorm2d(self.cfg.final_ch)
            init_weight(self.bn)
        # the reference stores the flag and then overwrites it with the module (models/generators.py:80-81)
        self.attention = Attention(G_ch * 2) if attention else False
        self.final = conv2d_lp(self.cfg.final_ch, img_ch)

        self._engine: Optional[Engine] = None
        self._engine_key = None
        self._seq: Optional[SequentialHalo] = None
        self._test_backend = None      # tests/ inject their launch emulator here to check the host logic without a GPU

    # ------------------------------------------------------------------------------------------
    def _weights_key(self):
        p = next(self.parameters())
        return (p.device, self.precision, tuple((t.data_ptr(), t._version) for t in self.state_dict().values()))

    def engine(self) -> Engine:
        """Pack the current parameters (once per parameter version) and return the launch engine."""
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            dev = next(self.parameters()).device
            if dev.type != "cuda" and self._test_backend is None:
                raise L.ItgError("ResidualPatchGenerator computes on CUDA devices only: move it with .to('cuda') "
                                 "(there is no CPU fallback on this path)")
            self._engine = Engine(self.cfg, self.state_dict(), self.precision, dev, backend=self._test_backend)
            self._engine_key = key
        return self._engine

    def reset_halo_state(self) -> None:
        """Forget the stored halos of the sequential protocol (a new Generator instance in the reference)."""
        self._seq = None

    @staticmethod
    def merge_map_crops(crops: torch.Tensor, nph: int, npw: int) -> torch.Tensor:
        """Inverse of utils.py:345-351: (nph*npw, 1, r+4, r+4) overlapping crops (stride r) -> (nph*r+4, npw*r+4)."""
        r = crops.shape[-1] - 4
        out = torch.empty((nph * r + 4, npw * r + 4), dtype=crops.dtype, device=crops.device)
        c = crops.reshape(nph, npw, r + 4, r + 4)
        for i in range(nph):
            for j in range(npw):
                out[i * r:i * r + r + 4, j * r:j * r + r + 4] = c[i, j]
        return out

    def forward(self, z, maps=None, image_location="1st_row_1st_col"):
        """z: (N, z_dim, nph*base_res+2, npw*base_res+2) merged latent grid with its random 1-px ring;
        maps (SSM): per level, either the reference's (N*nph*npw, 1, r+4, r+4) patch crops or the merged
        (N, 1, nph*r+4, npw*r+4) map.  Returns (N*nph*npw, img_ch, P, P) fp32 patches, row-major."""
        if self.training:
            raise NotImplementedError("training-mode forward (batch statistics, autograd) is out of scope: call .eval()")
        if self.padding_mode == "zeros":
            return self._forward_nonlocal(z, maps)
        nph, npw = LocalPadder.num_patches_h, LocalPadder.num_patches_w
        cfg, b = self.cfg, self.cfg.base_res
        if z.dim() != 4 or z.shape[1] != cfg.z_dim or z.shape[2] != nph * b + 2 or z.shape[3] != npw * b + 2:
            raise ValueError(f"z must be (N, {cfg.z_dim}, {nph * b + 2}, {npw * b + 2}) for a {nph}x{npw} patch grid, "
                             f"got {tuple(z.shape)}")
        eng = self.engine()
        N = z.shape[0]
        plan = eng.plan(nph, npw, L.IMG_PATCHES)
        seq_mode = not ("1st_row" in image_location and "1st_col" in image_location)
        if "1st_row" in image_location and "1st_col" in image_location:
            self._seq = SequentialHalo(eng.backend, nph, npw, cfg.border_is_replicate)
        elif self._seq is None:
            raise RuntimeError(f"image_location {image_location!r} without a preceding '1st_row_1st_col' call "
                               "(models/layers.py:86 would fail on a None halo)")
        if N > 1 and seq_mode:
            raise NotImplementedError("the sequential protocol is defined for one image (utils.py:341)")
        outs: List[torch.Tensor] = []
        for n in range(N):
            mp = None
            if cfg.type_norm == "SSM":
                if maps is None or maps[0] is None:
                    raise ValueError("type_norm='SSM' needs the per-level noise maps (utils.py:237-256)")
                mp = []
                for i in range(cfg.n_layers_G):
                    m = maps[i]
                    r = cfg.level_res(i + 1)
                    if m.shape[0] == N * nph * npw and m.shape[-1] == r + 4 and m.shape[-2] == r + 4:
                        m = self.merge_map_crops(m[n * nph * npw:(n + 1) * nph * npw], nph, npw)
                    else:
                        m = m[n, 0]
                    mp.append(m.float())
            with eng._on_device():                    # launches go to the current stream of the engine's device
                plan.set_inputs(z[n].float(), mp)
                plan.run(self._seq.hooks(plan, image_location) if N == 1 else None)
                outs.append(plan.out.clone())
        return outs[0] if N == 1 else torch.cat(outs, 0)

    def _forward_nonlocal(self, z, maps=None):
        """The non-local Generator (--padding_mode zeros, models/layers.py:19-27: every conv is conv3x3(..., p=1) on the whole image, no
        LocalPadder, image_location ignored): z (N, z_dim, h, w) -> (N, img_ch, h * 2^(n-1), w * 2^(n-1)); maps (SSM): per level
        (N, 1, h * 2^i, w * 2^i) (utils.sample_from_gen, utils.py:530-575).  Same kernels as the local path with a zero frame; the latent
        grid of h x w pixels plays the role of the patch grid.  Attention (over the whole level-3 map) is served for 8x8 / 16x16 maps."""
        cfg = self.cfg
        if z.dim() != 4 or z.shape[1] != cfg.z_dim:
            raise ValueError(f"z must be (N, {cfg.z_dim}, h, w), got {tuple(z.shape)}")
        N, _, h, w = z.shape
        eng = self.engine()
        plan = eng.plan(h, w, L.IMG_MERGED)
        outs: List[torch.Tensor] = []
        for n in range(N):
            mp = None
            if cfg.type_norm == "SSM":
                if maps is None or maps[0] is None:
                    raise ValueError("type_norm='SSM' needs the per-level noise maps (utils.py:558-564)")
                mp = [maps[i][n, 0].float() for i in range(cfg.n_layers_G)]
            with eng._on_device():
                plan.set_inputs(z[n].float(), mp)
                plan.run()
                outs.append(plan.out.clone())
        return outs[0] if N == 1 else torch.cat(outs, 0)
