"""Shared helpers of the test-suite: golden fixtures (tests/golden/*.npz, outputs of the unmodified reference)
and construction of the drop-in Generator from the same seeded weights."""
import ast
import glob
import os

import numpy as np
import torch

from oracle import itg_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "gen_*.npz")))


def load_case(name):
    """-> (npz, ctor kwargs, oracle cfg, state dict, z, maps)"""
    d = np.load(os.path.join(GOLD, name + ".npz"))
    kw = ast.literal_eval(str(d["cfg"]))
    ocfg = O.GenCfg(**kw)
    sd = O.make_state_dict(ocfg, int(d["wseed"]), stress=bool(d["stress"]))
    z = torch.from_numpy(d["z"])
    maps = None
    if ocfg.type_norm == "SSM":
        maps = [torch.from_numpy(d[f"map{i}"]) for i in range(ocfg.n_layers_G)]
    return d, kw, ocfg, sd, z, maps


def make_generator(kw, sd, precision, device=None, backend=None):
    import infinite_texture_gans_b200 as itg
    net = itg.ResidualPatchGenerator(**kw, precision=precision)
    net.load_state_dict(sd, strict=True)          # pins the reference's state-dict key / shape layout
    net.eval()
    if device is not None:
        net = net.to(device)
    net._test_backend = backend
    return net


def golden_image(d, key):
    """Full golden image, or None when the fixture only stores a strided sample (+ moments)."""
    return torch.from_numpy(d[key]) if key in d else None


def compare_with_golden(d, key, img: torch.Tensor, atol: float, H=None, W=None):
    img = img.detach().float().cpu()
    if key in d:
        ref = torch.from_numpy(d[key])
        if H is not None:
            ref = ref[:, :, :H, :W]
        assert tuple(img.shape) == tuple(ref.shape), (img.shape, ref.shape)
        err = (img - ref).abs().max().item()
    else:
        ref = torch.from_numpy(d[key + "_sub"])
        err = (img[:, :, ::3, ::3] - ref).abs().max().item()
    assert err <= atol, f"{key}: max-abs {err:.3e} > {atol:g}"
    return err


NONLOCAL_CASES = ("bn4_att_b4", "bn5_noatt_b11", "ssm4_noatt_b6", "bn4_noatt_tiles_b40")


def load_nonlocal_case(name):
    """tests/golden/nonlocal.npz (reference utils.sample_from_gen with padding_mode='zeros') -> (golden image, kwargs, oracle cfg, weights,
    z, maps, tiles): the noise is re-drawn in the reference's order (utils.py:549, 562) from the stored seed."""
    d = np.load(os.path.join(GOLD, "nonlocal.npz"))
    kw = ast.literal_eval(str(d[name + "_cfg"]))
    b, tiles, wseed, nseed = (int(v) for v in d[name + "_args"])
    ocfg = O.GenCfg(**kw)
    sd = O.make_state_dict(ocfg, wseed, stress=True)
    torch.manual_seed(nseed)
    z = torch.randn(1, kw["z_dim"], b, b)
    maps = None
    if kw["type_norm"] == "SSM":
        maps = [torch.randn(1, 1, b * 2 ** i, b * 2 ** i) for i in range(kw["n_layers_G"])]
    return torch.from_numpy(d[name + "_img"]), kw, ocfg, sd, z, maps, bool(tiles)
