"""Host-side mirror of the reference's models/layers.py surface for the inference path.

Only what callers of the Generator touch is kept: the class-level `LocalPadder` configuration
(models/layers.py:51-63), which `utils.sample_from_gen_PatchByPatch_train/_test` and the Generator
constructor mutate, and parameter containers with the reference's state-dict key layout.
The arithmetic of these layers lives in libitg_b200.so; none of the modules below has a forward.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class LocalPadder:
    """Class-level patch-grid configuration (models/layers.py:51-63).  The padding itself is the frame of
    the grid tensors (include/itg.h) and the conv kernels' addressing; there is no module to call."""
    num_patches_h = 3
    num_patches_w = 3
    outer_padding = "replicate"
    padding_size = 1
    conv_reduction = 2

    @classmethod
    def set_attributes(cls, num_patches_h=3, num_patches_w=3, outer_padding="replicate", padding_size=1,
                       conv_reduction=2):
        if padding_size != 1 or conv_reduction != 2:
            raise ValueError("only 3x3 convolutions with a 1-pixel local padding exist on this path "
                             "(padding_size=1, conv_reduction=2; test_sample.py:57-58)")
        cls.num_patches_h = num_patches_h
        cls.num_patches_w = num_patches_w
        cls.outer_padding = outer_padding
        cls.padding_size = padding_size
        cls.conv_reduction = conv_reduction


def init_weight(m: nn.Module) -> None:
    """Initialisation scheme of utils.py:745-762, selected by class NAME like the reference (`classname.find(...)`): orthogonal
    (gain 1) weights and zero biases for anything called *Conv* or *Linear*, weight ~ N(1, 0.02) and zero bias for *Batch*
    (affine norms only), orthogonal weights for *Embedding*.  Consumes the global RNG exactly like the reference, so the same
    seed gives the same "random-init weights" (pinned by tests/golden/aux.npz)."""
    name = m.__class__.__name__
    if name.find("Conv") != -1 or name.find("Linear") != -1:
        if getattr(m, "weight", None) is None:     # containers such as conv2d_lp: their nn.Conv2d child is visited by .apply()
            return
        nn.init.orthogonal_(m.weight, gain=1)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif name.find("Batch") != -1:
        if m.weight is not None:
            nn.init.normal_(m.weight, 1.0, 0.02)
            nn.init.zeros_(m.bias)
    elif name.find("Embedding") != -1:
        nn.init.orthogonal_(m.weight, gain=1)


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} only holds parameters; the Generator runs as fused CUDA launches "
                           "(infinite_texture_gans_b200.engine)")


class conv2d_lp(_NoForward):
    """Parameter container of conv2d_lp (models/layers.py:8-36): keys `conv.weight`, `conv.bias`."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=0)
        init_weight(self.conv)


class StochasticSpatialModulation(_NoForward):
    """Parameter container of StochasticSpatialModulation (models/layers.py:203-226): keys `bn.*`,
    `mlp_shared.0.*`, `embed.*`."""

    def __init__(self, in_channel, map_dim=1):
        super().__init__()
        self.bn = nn.BatchNorm2d(in_channel, affine=False)
        self.mlp_shared = nn.Sequential(nn.Conv2d(map_dim, 128, 3, padding=0), nn.ReLU())
        self.embed = nn.Conv2d(128, 2 * in_channel, 3, padding=0)
        init_weight(self.mlp_shared[0])
        init_weight(self.embed)
        with torch.no_grad():                       # models/layers.py:225-226
            nn.init.orthogonal_(self.embed.weight[:, :in_channel], gain=1)
            self.embed.weight[:, in_channel:].zero_()


class Attention(_NoForward):
    """Parameter container of Attention (models/layers.py:236-244)."""

    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.theta = nn.Conv2d(channels, channels // 8, 1)
        self.phi = nn.Conv2d(channels, channels // 8, 1)
        self.g = nn.Conv2d(channels, channels // 2, 1)
        self.o = nn.Conv2d(channels // 2, channels, 1)
        for m in (self.theta, self.phi, self.g, self.o):
            init_weight(m)
        self.gamma = nn.Parameter(torch.tensor(0.0))


class ResBlockGenerator(_NoForward):
    """Parameter container of ResBlockGenerator (models/layers.py:260-299)."""

    def __init__(self, in_channels, out_channels, type_norm="BN", map_dim=1):
        super().__init__()
        self.conv1 = conv2d_lp(in_channels, out_channels)
        self.conv2 = conv2d_lp(out_channels, out_channels)
        self.learnable_sc = in_channels != out_channels
        if self.learnable_sc:
            self.conv3 = nn.Conv2d(in_channels, out_channels, 1)
            init_weight(self.conv3)
        if type_norm == "SSM":
            self.bn1 = StochasticSpatialModulation(in_channels, map_dim)
            self.bn2 = StochasticSpatialModulation(out_channels, map_dim)
            if self.learnable_sc:
                self.bn3 = StochasticSpatialModulation(in_channels, map_dim)
        else:
            self.bn1 = nn.BatchNorm2d(in_channels)
            self.bn2 = nn.BatchNorm2d(out_channels)
            init_weight(self.bn1)
            init_weight(self.bn2)
