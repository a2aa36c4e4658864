"""Weight pre-packing: reference state_dict tensors -> the operand layouts of itg_conv_fwd.

Everything here is one-time host-side preparation (torch ops on whatever device the weights live on);
it is exercised on CPU by the test-suite.  Layouts are documented in include/itg.h.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .ops import c_store, k_pad_of, n_pad_of

BN_EPS = 1e-5   # nn.BatchNorm2d default used by the reference (models/layers.py:279-280)


def _pad_nk(w: torch.Tensor, n_pad: int, k_pad: int) -> torch.Tensor:
    """[taps, n, k] -> zero-padded [taps, n_pad, k_pad]."""
    t, n, k = w.shape
    out = torch.zeros(t, n_pad, k_pad, dtype=w.dtype, device=w.device)
    out[:, :n, :k] = w
    return out


def pack_conv3x3(weight: torch.Tensor, dtype, n_pad: Optional[int] = None) -> torch.Tensor:
    """(Cout, Cin, 3, 3) -> [9, n_pad, k_pad], tap t = ky*3+kx (cross-correlation, like F.conv2d)."""
    co, ci = weight.shape[:2]
    w = weight.float().permute(2, 3, 0, 1).reshape(9, co, ci)
    return _pad_nk(w, n_pad or n_pad_of(co), k_pad_of(c_store(ci))).to(dtype).contiguous()


def pack_conv1x1(weight: torch.Tensor, dtype) -> torch.Tensor:
    """(Cout, Cin, 1, 1) -> [1, n_pad, k_pad]."""
    co, ci = weight.shape[:2]
    w = weight.float().reshape(1, co, ci)
    return _pad_nk(w, n_pad_of(co), k_pad_of(c_store(ci))).to(dtype).contiguous()


# rows (or columns) of the 3x3 kernel that fall on low-res tap i of output parity a:
#   parity 0 (even output row 2y):   up-sampled rows 2y-1, 2y, 2y+1 -> low-res y-1 | y, y     -> {0}, {1,2}
#   parity 1 (odd output row 2y+1):  up-sampled rows 2y, 2y+1, 2y+2 -> low-res y, y | y+1     -> {0,1}, {2}
_PHASE_ROWS = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}


def pack_upconv(weight: torch.Tensor, dtype) -> torch.Tensor:
    """conv3x3(pad(nearest_up2(x))) as four 2x2 phase convs on the low-res framed tensor.

    (Cout, Cin, 3, 3) -> [16, n_pad, k_pad]; tap index = (a*2+b)*4 + i*2 + j reads low-res pixel
    (y + a-1+i, x + b-1+j) and contributes to output pixel (2y+a, 2x+b).  Sums are formed in fp32."""
    co, ci = weight.shape[:2]
    w = weight.float()
    taps = []
    for a in (0, 1):
        for b in (0, 1):
            for i in (0, 1):
                for j in (0, 1):
                    acc = torch.zeros(co, ci, dtype=torch.float32, device=w.device)
                    for r in _PHASE_ROWS[(a, i)]:
                        for s in _PHASE_ROWS[(b, j)]:
                            acc = acc + w[:, :, r, s]
                    taps.append(acc)
    return _pad_nk(torch.stack(taps, 0), n_pad_of(co), k_pad_of(c_store(ci))).to(dtype).contiguous()


def pad_vec(v: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros(n, dtype=torch.float32, device=v.device)
    out[: v.numel()] = v.float().reshape(-1)
    return out


def bn_fold(sd: Dict[str, torch.Tensor], prefix: str, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm2d as per-channel scale/shift (fp32, zero-padded to n)."""
    rstd = torch.rsqrt(sd[prefix + "running_var"].float() + BN_EPS)
    scale = sd[prefix + "weight"].float() * rstd
    shift = sd[prefix + "bias"].float() - sd[prefix + "running_mean"].float() * scale
    return pad_vec(scale, n), pad_vec(shift, n)


def bn_stats(sd: Dict[str, torch.Tensor], prefix: str, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Affine-free BN of StochasticSpatialModulation (layers.py:218): mean and 1/sqrt(var+eps), padded with zeros."""
    rstd = torch.rsqrt(sd[prefix + "running_var"].float() + BN_EPS)
    return pad_vec(sd[prefix + "running_mean"], n), pad_vec(rstd, n)


def pack_ssm_embed(weight: torch.Tensor, bias: torch.Tensor, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """SSM `embed` conv (2C, 128, 3, 3): interleave gamma/beta rows so that GEMM columns (2c, 2c+1) are
    (gamma_c, beta_c) of stored channel c (layers.py:232-233 chunk(2, dim=1))."""
    two_c = weight.shape[0]
    C = two_c // 2
    cs = c_store(C)
    w = torch.zeros(2 * cs, weight.shape[1], 3, 3, dtype=torch.float32, device=weight.device)
    b = torch.zeros(2 * cs, dtype=torch.float32, device=weight.device)
    w[0:2 * C:2] = weight[:C].float()
    w[1:2 * C:2] = weight[C:].float()
    b[0:2 * C:2] = bias[:C].float()
    b[1:2 * C:2] = bias[C:].float()
    n_pad = n_pad_of(2 * cs)
    return pack_conv3x3(w, dtype, n_pad=n_pad), pad_vec(b, n_pad)


def pack_ssm_mlp(weight: torch.Tensor, bias: torch.Tensor, dtype) -> torch.Tensor:
    """SSM `mlp_shared` conv (128, 1, 3, 3) + bias -> the [128, 16] K-major operand of itg_ssm_fwd: columns 0..8 are the taps
    (ky*3+kx), columns 9 and 10 carry the bias as a 16-bit hi + lo pair (the kernel multiplies them by constant ones, so the
    fp32 accumulator receives the bias to ~2^-22 relative), the rest is zero (layers.py:220-222, 229)."""
    n = weight.shape[0]
    if tuple(weight.shape[1:]) != (1, 3, 3):
        raise ValueError("the fused SSM kernel serves map_dim = 1 (test_sample.py:56 passes map_dim=1)")
    out = torch.zeros(n, 16, dtype=torch.float32, device=weight.device)
    out[:, :9] = weight.float().reshape(n, 9)
    hi = bias.float().to(dtype).float()
    out[:, 9] = hi
    out[:, 10] = (bias.float() - hi).to(dtype).float()
    return out.to(dtype).contiguous()
