#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Every fixture stores: the constructor config, the weight seed (weights are regenerated with
oracle.itg_oracle.make_state_dict and were loaded into the reference Generator with
load_state_dict(strict=True), which also pins the state-dict key/shape layout), the noise seed,
and the reference's outputs.  Nothing from the reference's sources is copied; only its outputs are.

Fixtures:
  gen_*.npz    -- full sampler / one-shot Generator outputs of the reference:
                  'seq'  = utils.sample_from_gen_PatchByPatch_test (utils.py:258-397), 3x3 sub-image stepping
                  'one'  = one forward with LocalPadder.set_attributes(total_h,total_w) (the train-time call,
                           utils.py:475-527), merged with utils.merge_patches_into_image
  nonlocal.npz -- utils.sample_from_gen outputs of the non-local Generator (padding_mode='zeros'), plain and with tile_process.
  aux.npz      -- utils.build_z / utils.build_maps outputs (utils.py:221-256) and utils.init_weight results (utils.py:745-762)
                  under fixed seeds: pins the host-side noise plumbing and the "random-init weights" scheme.
  localpad.npz -- integer-coded tensors pushed through models.layers.LocalPadder in eval mode for a whole
                  sequential sweep (all location classes) and both outer paddings: pins the halo indexing.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import itg_oracle as O  # noqa: E402

import utils as ref_utils  # noqa: E402  (reference)
from models import generators as ref_gen  # noqa: E402  (reference)
from models import layers as ref_layers  # noqa: E402  (reference)

torch.set_num_threads(8)

CASES = {
    # name: (cfg kwargs, H, W, weight seed, noise seed, stress)
    "gen_bn4_att_rep": (dict(z_dim=16, G_ch=8, n_layers_G=4, attention=True, leak=0.02, type_norm="BN",
                             outer_padding="replicate"), 160, 160, 11, 12, True),
    "gen_bn4_noatt_const_crop": (dict(z_dim=16, G_ch=8, n_layers_G=4, attention=False, leak=0.0, type_norm="BN",
                                      outer_padding="constant"), 80, 200, 21, 22, True),
    "gen_ssm4_att_rep": (dict(z_dim=16, G_ch=8, n_layers_G=4, attention=True, leak=0.02, type_norm="SSM",
                              outer_padding="replicate"), 96, 160, 31, 32, True),
    "gen_bn5_gamma0_rep": (dict(z_dim=16, G_ch=8, n_layers_G=5, attention=True, leak=0.02, type_norm="BN",
                                outer_padding="replicate"), 192, 320, 41, 42, False),
    "gen_bn6_stock_rep": (dict(z_dim=16, G_ch=8, n_layers_G=6, attention=True, leak=0.02, type_norm="BN",
                               outer_padding="replicate"), 384, 384, 51, 52, False),
    "gen_241_3x3": (dict(z_dim=128, G_ch=52, n_layers_G=6, attention=True, leak=0.02, type_norm="BN",
                         outer_padding="replicate"), 384, 384, 61, 62, False),
}


def build_ref(cfg: O.GenCfg, sd):
    net = ref_gen.ResidualPatchGenerator(
        z_dim=cfg.z_dim, G_ch=cfg.G_ch, base_res=cfg.base_res, n_layers_G=cfg.n_layers_G, attention=cfg.attention,
        img_ch=cfg.img_ch, leak=cfg.leak, SN=False, type_norm=cfg.type_norm, map_dim=1, padding_mode="local",
        outer_padding=cfg.outer_padding, num_patches_h=3, num_patches_w=3, padding_size=1, conv_reduction=2)
    net.load_state_dict(sd, strict=True)
    return net.eval()


def draw_noise_like_reference(cfg, total_h, total_w, seed):
    """Same global-RNG draw order as utils.py:228 then :246 -> the reference sampler sees this noise."""
    torch.manual_seed(seed)
    b = cfg.base_res
    z = torch.randn(1, cfg.z_dim, total_h * b + 2, total_w * b + 2)
    maps = None
    if cfg.type_norm == "SSM":
        maps = [torch.randn(1, 1, total_h * b * 2 ** i + 4, total_w * b * 2 ** i + 4) for i in range(cfg.n_layers_G)]
    return z, maps


def gen_case(name, kw, H, W, wseed, nseed, stress):
    cfg = O.GenCfg(**kw)
    sd = O.make_state_dict(cfg, wseed, stress=stress)
    geo = O.geometry(H, W, cfg)
    th, tw = geo["total_h"], geo["total_w"]
    z, maps = draw_noise_like_reference(cfg, th, tw, nseed)

    # (B) the shipped sequential sampler
    net = build_ref(cfg, sd)
    torch.manual_seed(nseed)
    with torch.no_grad():
        seq = ref_utils.sample_from_gen_PatchByPatch_test(
            net, z_dim=cfg.z_dim, base_res=cfg.base_res, num_images=1, device="cpu",
            output_resolution_height=H, output_resolution_width=W)

    # (A) one-shot forward on the full grid, the way sample_from_gen_PatchByPatch_train calls it
    ref_layers.LocalPadder.set_attributes(num_patches_h=th, num_patches_w=tw, outer_padding=cfg.outer_padding,
                                          padding_size=1, conv_reduction=2)
    net1 = build_ref(cfg, sd)
    ref_layers.LocalPadder.set_attributes(num_patches_h=th, num_patches_w=tw, outer_padding=cfg.outer_padding,
                                          padding_size=1, conv_reduction=2)
    maps_p = None
    if maps is not None:
        maps_p = [ref_utils.crop_images(m, cfg.base_res * 2 ** i + 4, cfg.base_res * 2 ** i + 4, cfg.base_res * 2 ** i)
                  for i, m in enumerate(maps)]
    with torch.no_grad():
        patches = net1(z, maps_p, "1st_row_1st_col")
        one = ref_utils.merge_patches_into_image(patches, th, tw)
    ref_layers.LocalPadder.set_attributes()  # back to defaults

    out = dict(cfg=np.array(repr(kw)), H=H, W=W, wseed=wseed, nseed=nseed, stress=stress,
               total_h=th, total_w=tw, z=z.numpy(), seq=seq.numpy().astype(np.float32),
               one=one.numpy().astype(np.float32), patches_shape=np.array(patches.shape))
    if maps is not None:
        for i, m in enumerate(maps):
            out[f"map{i}"] = m.numpy()
    if name == "gen_241_3x3":
        # keep the big-channel case small: store a strided sample of the image plus its moments
        for k in ("seq", "one"):
            full = out.pop(k)
            out[k + "_sub"] = full[:, :, ::3, ::3].copy()
            out[k + "_sum"] = np.float64(full.astype(np.float64).sum())
            out[k + "_abs"] = np.float64(np.abs(full.astype(np.float64)).sum())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    d = float((seq[:, :, :H, :W] - one[:, :, :H, :W]).abs().max())
    print(f"{name}: grid {th}x{tw} out {tuple(seq.shape)}  max|seq-one| = {d:.3e}")


def localpad_case():
    """Integer-coded merged inputs through the reference LocalPadder for a 3x4-step sweep."""
    out = {}
    for mode in ("replicate", "constant"):
        ref_layers.LocalPadder.set_attributes(num_patches_h=3, num_patches_w=3, outer_padding=mode,
                                              padding_size=1, conv_reduction=2)
        pad = ref_layers.LocalPadder(True).eval()
        steps_h, steps_w, r, C = 3, 4, 4, 2
        k = 0
        for ih in range(steps_h):
            for iw in range(steps_w):
                loc = O.image_location(ih, iw, steps_h, steps_w)
                # value encodes (step, channel, patch, row, col) uniquely
                x = (torch.arange(9 * C * r * r, dtype=torch.float32).reshape(9, C, r, r) + 1) + 1000.0 * (k + 1)
                y = pad(x, loc)
                out[f"{mode}_{k}_in"] = x.numpy()
                out[f"{mode}_{k}_out"] = y.numpy()
                out[f"{mode}_{k}_loc"] = np.array(loc)
                k += 1
    ref_layers.LocalPadder.set_attributes()
    np.savez_compressed(os.path.join(HERE, "localpad.npz"), **out)
    print("localpad: done")


def nonlocal_cases():
    """The non-local Generator (padding_mode='zeros') through utils.sample_from_gen (utils.py:530-575), plain and tiled
    (utils.tile_process, utils.py:401-470), and test_sample.py's base_res arithmetic (test_sample.py:70-73)."""
    out = {}
    cases = {
        # name: (cfg kwargs, base_res passed to sample_from_gen, tiles, weight seed, noise seed)
        "bn4_att_b4": (dict(z_dim=16, G_ch=8, n_layers_G=4, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate"), 4, False, 81, 82),
        "bn5_noatt_b11": (dict(z_dim=16, G_ch=8, n_layers_G=5, attention=False, leak=0.0, type_norm="BN", outer_padding="replicate"), 11, False, 83, 84),
        "ssm4_noatt_b6": (dict(z_dim=16, G_ch=8, n_layers_G=4, attention=False, leak=0.02, type_norm="SSM", outer_padding="replicate"), 6, False, 85, 86),
        "bn4_noatt_tiles_b40": (dict(z_dim=16, G_ch=8, n_layers_G=4, attention=False, leak=0.02, type_norm="BN", outer_padding="replicate"), 40, True, 87, 88),
    }
    for name, (kw, b, tiles, wseed, nseed) in cases.items():
        cfg = O.GenCfg(**kw)
        sd = O.make_state_dict(cfg, wseed, stress=True)
        net = ref_gen.ResidualPatchGenerator(
            z_dim=cfg.z_dim, G_ch=cfg.G_ch, base_res=cfg.base_res, n_layers_G=cfg.n_layers_G, attention=cfg.attention, img_ch=cfg.img_ch,
            leak=cfg.leak, SN=False, type_norm=cfg.type_norm, map_dim=1, padding_mode="zeros", outer_padding=cfg.outer_padding)
        net.load_state_dict(sd, strict=True)
        net.eval()
        torch.manual_seed(nseed)
        with torch.no_grad():
            img = ref_utils.sample_from_gen(net, z_dim=cfg.z_dim, base_res=b, num_images=1, tiles=tiles, device="cpu")
        out[name + "_cfg"] = np.array(repr(kw))
        out[name + "_args"] = np.array([b, int(tiles), wseed, nseed])
        out[name + "_img"] = img.numpy().astype(np.float32)
        print(f"nonlocal {name}: {tuple(img.shape)}")
    np.savez_compressed(os.path.join(HERE, "nonlocal.npz"), **out)


def aux_case():
    """build_z / build_maps (noise draw + overlapping sub-image crops) and init_weight, straight from the reference."""
    import torch.nn as nn
    out = {}
    torch.manual_seed(71)
    z = ref_utils.build_z(num_images=1, z_dim=6, base_res=4, num_patches_height=3, num_patches_width=3,
                          total_num_patches_height=5, total_num_patches_width=7)
    out["build_z"] = z.numpy()
    torch.manual_seed(72)
    maps = ref_utils.build_maps(num_images=1, map_dim=1, n_layers_G=3, base_res=4, num_patches_height=3, num_patches_width=3,
                                total_num_patches_height=5, total_num_patches_width=7)
    for i, m in enumerate(maps):
        out[f"build_maps{i}"] = m.numpy()
    torch.manual_seed(73)
    mods = dict(conv3=nn.Conv2d(5, 7, 3), conv1=nn.Conv2d(12, 4, 1), lin=nn.Linear(6, 9), bn=nn.BatchNorm2d(10), emb=nn.Embedding(11, 8))
    for m in mods.values():
        m.apply(ref_utils.init_weight)
    for k, m in mods.items():
        out[f"init_{k}_weight"] = m.weight.detach().numpy()
        if getattr(m, "bias", None) is not None:
            out[f"init_{k}_bias"] = m.bias.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "aux.npz"), **out)
    print("aux: done", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, args in CASES.items():
        if only and name not in only:
            continue
        gen_case(name, *args)
    if not only or "localpad" in only:
        localpad_case()
    if not only or "aux" in only:
        aux_case()
    if not only or "nonlocal" in only:
        nonlocal_cases()
