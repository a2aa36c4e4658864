"""CPU tests of the host side: launch plans (engine.py), weight packing (packing.py), the sequential halo state
machine (halo.py), the sampler geometry (utils.py) and the C-ABI symbol table.  The CUDA kernels are replaced by
tests/emulator.py (fp32 torch restatement of each launch); results are held to the reference's golden outputs."""
import ctypes
import os
import re

import pytest
import torch
from hypothesis import given, settings, strategies as st

import infinite_texture_gans_b200 as itg
from common import CASES, compare_with_golden, load_case, make_generator
from emulator import EmulatorBackend
from infinite_texture_gans_b200 import _lib as L
from infinite_texture_gans_b200.config import GenConfig, flops_per_patch
from oracle import itg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", CASES)
def test_plan_oneshot_matches_reference(name):
    d, kw, ocfg, sd, z, maps = load_case(name)
    net = make_generator(kw, sd, "fp32", backend=EmulatorBackend())
    img = itg.utils.generate_full_grid(net, z, maps)
    compare_with_golden(d, "one", img, 5e-5)


@pytest.mark.parametrize("name", [c for c in CASES if "241" not in c])
def test_sequential_protocol_matches_reference(name):
    """image_location state machine on device buffers == LocalPadder's (layers.py:78-143), gamma != 0 included."""
    d, kw, ocfg, sd, z, maps = load_case(name)
    net = make_generator(kw, sd, "fp32", backend=EmulatorBackend())
    H, W = int(d["H"]), int(d["W"])
    img = itg.utils.sample_from_gen_PatchByPatch_test(net, z_dim=kw["z_dim"], output_resolution_height=H,
                                                      output_resolution_width=W, schedule="sequential", noise=(z, maps))
    compare_with_golden(d, "seq", img, 5e-5)


def test_reference_style_map_crops_are_accepted():
    """forward() takes the per-patch (nph*npw,1,r+4,r+4) map crops the reference sampler builds (utils.py:345-351)."""
    d, kw, ocfg, sd, z, maps = load_case("gen_ssm4_att_rep")
    th, tw = int(d["total_h"]), int(d["total_w"])
    net = make_generator(kw, sd, "fp32", backend=EmulatorBackend())
    itg.LocalPadder.set_attributes(num_patches_h=th, num_patches_w=tw)
    try:
        crops = [itg.utils.crop_images(m, 4 * 2 ** i + 4, 4 * 2 ** i + 4, 4 * 2 ** i) for i, m in enumerate(maps)]
        patches = net(z, crops)
    finally:
        itg.LocalPadder.set_attributes()
    compare_with_golden(d, "one", itg.utils.merge_patches_into_image(patches, th, tw), 5e-5)


def test_fused_ssm_plan_matches_reference():
    """16-bit plans run every StochasticSpatialModulation as ONE itg_ssm_fwd launch (hidden map on chip): the launch list has no
    tap-stack / mlp / embed launches, needs a fraction of the two-launch plan's memory, and (emulated) still reproduces the reference."""
    from infinite_texture_gans_b200.engine import Engine, Plan
    d, kw, ocfg, sd, z, maps = load_case("gen_ssm4_att_rep")
    th, tw = int(d["total_h"]), int(d["total_w"])
    eng = Engine(GenConfig(**kw), sd, "fp16", "cpu", backend=EmulatorBackend())
    fused = Plan(eng.cfg, eng.weights, eng.backend, th, tw, eng.device, eng.impl)
    split = Plan(eng.cfg, eng.weights, eng.backend, th, tw, eng.device, eng.impl, fuse_ssm=False)
    kinds = [k for k, _ in fused.ops]
    n_norms = sum(3 if ci != co else 2 for ci, co in eng.cfg.block_channels())
    assert fused.fuse_ssm and kinds.count("ssm") == n_norms and "pack_map" not in kinds
    assert [k for k, _ in split.ops].count("pack_map") == eng.cfg.n_layers_G
    n_fold = sum(1 for k, o in fused.ops if k == "conv" and o.in2 is not None)        # 1x1 shortcuts folded into their conv2 (second input)
    assert n_fold == sum(1 for ci, co in eng.cfg.block_channels() if ci != co) and all(o.in2 is None for k, o in split.ops if k == "conv")
    assert fused.n_launches == split.n_launches - n_norms - eng.cfg.n_layers_G - n_fold
    assert fused.arena_bytes < 0.5 * split.arena_bytes
    outs = []
    for p in (fused, split):
        p.set_inputs(z, [m[0, 0] for m in maps])
        outs.append(p.run().clone())
        compare_with_golden(d, "one", outs[-1], 2e-2)
    assert (outs[0] - outs[1]).abs().max().item() < 1e-2     # same arithmetic up to the bias's hi + lo split; rounding flips of the 16-bit hidden map propagate


def test_inter_location_without_state_raises():
    d, kw, ocfg, sd, z, maps = load_case("gen_bn4_att_rep")
    net = make_generator(kw, sd, "fp32", backend=EmulatorBackend())
    with pytest.raises(RuntimeError, match="1st_row_1st_col"):
        net(torch.zeros(1, kw["z_dim"], 14, 14), None, "inter_row_inter_col")


def test_seeded_sampler_draws_reference_noise():
    """Same torch seed -> same z / maps as the reference draw order (utils.py:228, 246)."""
    torch.manual_seed(5)
    z, maps = itg.utils.draw_noise(1, 16, 4, 4, 1, "SSM", 3, 5)
    torch.manual_seed(5)
    z2 = torch.randn(1, 16, 14, 22)
    m2 = [torch.randn(1, 1, 3 * 4 * 2 ** i + 4, 5 * 4 * 2 ** i + 4) for i in range(4)]
    assert torch.equal(z, z2) and all(torch.equal(a, b) for a, b in zip(maps, m2))


@settings(max_examples=200, deadline=None)
@given(st.integers(33, 3000), st.integers(33, 3000), st.sampled_from([4, 5, 6]), st.sampled_from([2, 3, 4]))
def test_geometry_matches_oracle(H, W, n, nps):
    cfg = O.GenCfg(n_layers_G=n, num_patches_h=nps, num_patches_w=nps)
    if H <= cfg.patch_px or W <= cfg.patch_px:
        return
    g = itg.utils.patch_grid_geometry(H, W, n, 4, nps, nps)
    assert g == O.geometry(H, W, cfg)
    assert g["total_h"] * g["P"] >= H and g["total_w"] * g["P"] >= W


def test_crop_and_merge_match_oracle():
    x = torch.arange(2 * 3 * 14 * 22, dtype=torch.float32).reshape(2, 3, 14, 22)
    assert torch.equal(itg.utils.crop_images(x, 6, 6, 4), O.crop_windows(x, 6, 6, 4))
    p = torch.arange(12 * 2 * 4 * 4, dtype=torch.float32).reshape(12, 2, 4, 4)
    assert torch.equal(itg.utils.merge_patches_into_image(p, 2, 3), O.merge_patches(p, 2, 3))


def test_flop_model_matches_survey():
    f = lambda **k: flops_per_patch(GenConfig(z_dim=128, G_ch=52, **k)) / 1e6
    assert abs(f(n_layers_G=6) - 938.55) < 0.01
    assert abs(f(n_layers_G=4) - 608.73) < 0.01
    assert abs(f(n_layers_G=5, type_norm="SSM") - 5475.37) < 0.01


def test_buffer_reuse_does_not_change_results():
    from infinite_texture_gans_b200.engine import Engine, Plan
    d, kw, ocfg, sd, z, maps = load_case("gen_ssm4_att_rep")
    th, tw = int(d["total_h"]), int(d["total_w"])
    eng = Engine(GenConfig(**kw), sd, "fp32", "cpu", backend=EmulatorBackend())
    outs = []
    for reuse in (True, False):
        p = Plan(eng.cfg, eng.weights, eng.backend, th, tw, eng.device, eng.impl, reuse_buffers=reuse)
        p.set_inputs(z, [m[0, 0] for m in maps])
        outs.append((p.run().clone(), p.arena_bytes))
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][1] < outs[1][1]


def test_library_exports_every_declared_symbol():
    """include/itg.h <-> libitg_b200.so <-> the ctypes binding agree (no compute calls: runs without a GPU)."""
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    header = open(os.path.join(ROOT, "include", "itg.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(itg_\w+)\s*\(", header, flags=re.M))
    assert declared == set(L.EXPORTS)
    lib = ctypes.CDLL(L.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    lib2 = L.load()
    assert lib2.itg_version() == 3
    assert lib2.itg_conv_desc_size() == ctypes.sizeof(L.ConvDesc)
    assert lib2.itg_ssm_desc_size() == ctypes.sizeof(L.SsmDesc)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libitg_b200.so")
    with pytest.raises(L.ItgError, match="no CPU or PyTorch fallback"):
        L.load()


def test_output_stage_and_streaming_reject_host_tensors():
    """The device-side output stage and the streaming sampler validate their arguments before touching the library and have no host path."""
    import infinite_texture_gans_b200 as itg
    with pytest.raises(ValueError, match="CUDA"):
        itg.utils.image_to_uint8(torch.zeros(1, 3, 8, 8))
    with pytest.raises(ValueError, match="fp32"):
        itg.utils.image_to_uint8(torch.zeros(1, 3, 8, 8, dtype=torch.float16))
    d, kw, ocfg, sd, z, maps = load_case("gen_bn4_att_rep")
    net = make_generator(kw, sd, "fp16")                     # left on the CPU
    with pytest.raises(ValueError, match="out_format"):
        next(itg.utils.generate_textures(net, [(z, maps)], 64, 64, out_format="png"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "infinite_texture_gans_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("Oracle", ""), f


def test_cli_keeps_the_reference_flags():
    """test_sample.py:14-18 flag names and defaults."""
    from infinite_texture_gans_b200.cli import build_parser, load_G
    a = build_parser().parse_args([])
    assert (a.output_resolution_height, a.output_resolution_width, a.output_name, a.tiles) == (384, 384, "241_generated.jpg", False)
    assert a.model_path == "results/241_lp_bn_outerpadRepl/300__ema.pth"
    d, kw, ocfg, sd, z, maps = load_case("gen_bn4_att_rep")
    net = itg.ResidualPatchGenerator(**kw)
    load_G({"module." + k: v for k, v in sd.items()}, net)          # DataParallel checkpoints
    assert not net.training
    assert torch.equal(net.state_dict()["start.conv.weight"], sd["start.conv.weight"])


@pytest.mark.parametrize("nps,H,W", [(4, 224, 320), (5, 288, 416), (3, 150, 200)])
def test_sequential_protocol_other_subimage_sizes(nps, H, W):
    """The reference sampler takes num_patches_height/width (utils.py:258-259); square sub-images of 3..5 patches,
    sizes that are not multiples of the patch, gamma != 0: equal to the oracle's restatement of the shipped schedule."""
    kw = dict(z_dim=16, G_ch=8, n_layers_G=4, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate")
    ocfg = O.GenCfg(**kw, num_patches_h=nps, num_patches_w=nps)
    sd = O.make_state_dict(ocfg, 5, stress=True)
    geo = O.geometry(H, W, ocfg)
    z, _ = O.make_noise(ocfg, geo["total_h"], geo["total_w"], 3)
    with torch.no_grad():
        ref = O.sample_patch_by_patch(sd, ocfg, H, W, z)
    net = make_generator(kw, sd, "fp32", backend=EmulatorBackend())
    got = itg.utils.sample_from_gen_PatchByPatch_test(net, z_dim=16, num_patches_height=nps, num_patches_width=nps,
                                                      output_resolution_height=H, output_resolution_width=W,
                                                      schedule="sequential", noise=(z, None))
    assert tuple(got.shape) == (1, 3, H, W)
    assert (got - ref).abs().max().item() <= 5e-5


@pytest.mark.parametrize("name,nps", [("gen_bn4_att_rep", 3), ("gen_ssm4_att_rep", 3), ("gen_bn4_att_rep", 4)])
def test_row_stream_equals_the_sequential_sampler(name, nps):
    """utils.stream_texture_rows yields the texture band by band with O(sub-image row) state; fed the rows of a full noise grid it is the
    shipped schedule itself: the concatenated bands equal sample_from_gen_PatchByPatch_test(schedule='sequential') bit for bit, and an
    endless stream's first bands equal those of a longer texture (no 'last_row' is ever taken)."""
    d, kw, ocfg, sd, _, _ = load_case(name)
    net = make_generator(kw, sd, "fp32", backend=EmulatorBackend())
    P, b = net.cfg.patch_px, net.cfg.base_res
    H, W = (2 * (nps - 1) + 1) * P - 7, (2 * (nps - 1) + 1) * P - 3           # two sub-image steps each way, not multiples of the patch
    geo = itg.utils.patch_grid_geometry(H, W, kw["n_layers_G"], b, nps, nps)
    torch.manual_seed(11)
    z, maps = itg.utils.draw_noise(1, kw["z_dim"], b, kw["n_layers_G"], 1, kw["type_norm"], geo["total_h"], geo["total_w"])
    ref = itg.utils.sample_from_gen_PatchByPatch_test(net, z_dim=kw["z_dim"], num_patches_height=nps, num_patches_width=nps,
                                                      output_resolution_height=H, output_resolution_width=W, schedule="sequential", noise=(z, maps))

    def rows(ih):
        py = ih * (nps - 1)
        zr = z[:, :, py * b:(py + nps) * b + 2]
        mr = None if maps is None else [m[:, :, py * b * 2 ** i:(py + nps) * b * 2 ** i + 4] for i, m in enumerate(maps)]
        return zr, mr

    bands = list(itg.utils.stream_texture_rows(net, W, noise_rows=rows, steps=geo["steps_h"], num_patches_height=nps, num_patches_width=nps))
    assert [t.shape[2] for t in bands] == [(nps - 1) * P] * (geo["steps_h"] - 1) + [nps * P]
    got = torch.cat(bands, 2)[:, :, :H]
    assert torch.equal(got.cpu(), ref)
    endless = itg.utils.stream_texture_rows(net, W, noise_rows=rows, num_patches_height=nps, num_patches_width=nps)
    first = next(endless)
    endless.close()
    assert torch.equal(first.cpu(), ref[:, :, :(nps - 1) * P])
    assert (itg.LocalPadder.num_patches_h, itg.LocalPadder.num_patches_w) != (None, None)
    with pytest.raises(ValueError, match="either seed"):
        next(itg.utils.stream_texture_rows(net, W))


def test_build_z_build_maps_and_init_weight_match_the_reference():
    """utils.build_z / utils.build_maps (utils.py:221-256) and layers.init_weight (utils.py:745-762) against outputs of the unmodified
    reference under the same torch seeds (tests/golden/aux.npz, written by make_golden.py): same draw order, same overlapping
    sub-image crops (stride (npw-1)*res, row-major), same initialisation scheme."""
    import numpy as np
    import torch.nn as nn
    from common import GOLD
    d = np.load(os.path.join(GOLD, "aux.npz"))
    torch.manual_seed(71)
    z = itg.utils.build_z(num_images=1, z_dim=6, base_res=4, num_patches_height=3, num_patches_width=3,
                          total_num_patches_height=5, total_num_patches_width=7)
    assert z.dtype == torch.float32 and torch.equal(z, torch.from_numpy(d["build_z"]))
    torch.manual_seed(72)
    maps = itg.utils.build_maps(num_images=1, map_dim=1, n_layers_G=3, base_res=4, num_patches_height=3, num_patches_width=3,
                                total_num_patches_height=5, total_num_patches_width=7)
    assert len(maps) == 3
    for i, m in enumerate(maps):
        assert torch.equal(m, torch.from_numpy(d[f"build_maps{i}"])), i
    torch.manual_seed(73)
    mods = dict(conv3=nn.Conv2d(5, 7, 3), conv1=nn.Conv2d(12, 4, 1), lin=nn.Linear(6, 9), bn=nn.BatchNorm2d(10), emb=nn.Embedding(11, 8))
    for m in mods.values():
        m.apply(itg.layers.init_weight)
    for k, m in mods.items():
        assert torch.equal(m.weight.detach(), torch.from_numpy(d[f"init_{k}_weight"])), k
        if getattr(m, "bias", None) is not None:
            assert torch.equal(m.bias.detach(), torch.from_numpy(d[f"init_{k}_bias"])), k
    # properties the scheme guarantees (what BASELINE.json calls "random-init weights"): orthonormal rows or columns, zero biases
    w = mods["conv3"].weight.detach().reshape(7, -1)
    assert torch.allclose(w @ w.t(), torch.eye(7), atol=1e-5)
    assert float(mods["conv3"].bias.detach().abs().max()) == 0.0 and abs(float(mods["bn"].weight.detach().mean()) - 1.0) < 0.05


@pytest.mark.parametrize("name,precision,tol", [("bn4_att_b4", "fp32", 5e-5), ("bn5_noatt_b11", "fp32", 5e-5), ("bn4_noatt_tiles_b40", "fp32", 5e-5),
                                                ("ssm4_noatt_b6", "fp16", 2e-2)])
def test_nonlocal_generator_and_sampler(name, precision, tol):
    """--padding_mode zeros (the non-local Generator, test_sample.py:70-73): same launch plans with a zero frame; utils.sample_from_gen
    and utils.tile_process mirror utils.py:401-470,530-575.  Checked (emulated launches) against the reference's golden outputs."""
    from common import load_nonlocal_case
    gold, kw, ocfg, sd, z, maps, tiles = load_nonlocal_case(name)
    net = itg.ResidualPatchGenerator(**kw, padding_mode="zeros", precision=precision)
    net.load_state_dict(sd, strict=True)
    net.eval()
    net._test_backend = EmulatorBackend()
    img = itg.utils.sample_from_gen(net, z_dim=kw["z_dim"], base_res=z.shape[-1], tiles=tiles, noise=(z, maps))
    assert tuple(img.shape) == tuple(gold.shape)
    assert (img.float() - gold).abs().max().item() <= tol


def test_nonlocal_limits_are_reported():
    from common import load_nonlocal_case
    gold, kw, ocfg, sd, z, maps, tiles = load_nonlocal_case("bn4_att_b4")
    net = itg.ResidualPatchGenerator(**kw, padding_mode="zeros", precision="fp32")
    net.load_state_dict(sd)
    net.eval()
    net._test_backend = EmulatorBackend()
    with pytest.raises(NotImplementedError, match="attention"):
        net(torch.zeros(1, kw["z_dim"], 6, 6))                       # attention over a 24x24 map is not served
    with pytest.raises(ValueError, match="padding_mode"):
        itg.ResidualPatchGenerator(**kw, padding_mode="reflect")


def test_bf16_mode_is_refused_with_a_reason():
    """Single-pass bf16 operands miss the 2e-2 image tolerance; the product offers fp16 (same tensor-core rate) and says so."""
    d, kw, ocfg, sd, z, maps = load_case("gen_bn4_att_rep")
    net = make_generator(kw, sd, "bf16", backend=EmulatorBackend())
    with pytest.raises(ValueError, match="cannot meet the 2e-2"):
        net.engine()
