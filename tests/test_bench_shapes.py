"""Parity at the shapes bench.py times (BASELINE.json configs), full channel width.

cfg1 (9x9) and cfg2 (7x21) run whole through the CPU oracle (Oracle A, one-shot), pre- and post-tanh.  cfg3 (61x61 SSM), cfg4
(129x129) and the cfg5 row band (16x129) are too large for the oracle: the CUDA path computes the whole grid, the oracle recomputes
windows (a corner, an edge, the interior, the opposite corner) on the cropped noise with a two-patch margin
(oracle/window.py).  Weights: the reference's init scheme ("random-init weights", BASELINE.json) and the stress initialisation
(non-zero biases / running statistics, attention.gamma = 0.5).

Tolerances (north_star): fp32 mode <= 1e-3, 16-bit mode <= 2e-2 on the [-1, 1] image; measured values are printed.
"""
import pytest
import torch

from common import make_generator
from oracle import itg_oracle as O
from oracle import window as OW

pytestmark = pytest.mark.gpu

K241 = dict(z_dim=128, G_ch=52, n_layers_G=6, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate")
K34 = dict(z_dim=128, G_ch=52, n_layers_G=5, attention=True, leak=0.02, type_norm="SSM", outer_padding="replicate")
K417 = dict(z_dim=128, G_ch=52, n_layers_G=4, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate")
TOL = {"fp32": 1e-3, "fp16": 2e-2}
# pre-tanh values are not confined to [-1, 1]: the tolerance scales with the largest pre-activation (printed)
PRE_REL = {"fp32": 1e-3, "fp16": 2e-2}


@pytest.mark.parametrize("stress", [False, True], ids=["stock-init", "stress-init"])
@pytest.mark.parametrize("name,th,tw", [("cfg1", 9, 9), ("cfg2", 7, 21)])
def test_241_full_grid_vs_oracle(name, th, tw, stress):
    import infinite_texture_gans_b200 as itg
    ocfg = O.GenCfg(**K241)
    sd = O.make_state_dict(ocfg, seed=201, stress=stress, orthogonal=not stress)
    z, _ = O.make_noise(ocfg, th, tw, seed=202)
    with torch.no_grad():
        pre = O.forward_merged(sd, ocfg, z, pre_tanh=True)
    ref = torch.tanh(pre)
    scale = max(1.0, pre.abs().max().item())
    for precision in ("fp32", "fp16"):
        net = make_generator(K241, sd, precision, "cuda")
        img = itg.utils.generate_full_grid(net, z).cpu()
        gpre = itg.utils.generate_full_grid(net, z, pre_tanh=True).cpu()
        err, perr = (img - ref).abs().max().item(), (gpre - pre).abs().max().item()
        print(f"{name} {th}x{tw} {'stress' if stress else 'stock'} {precision}: post-tanh max-abs {err:.3e}, pre-tanh {perr:.3e} (max |pre| {scale:.2f})")
        assert err <= TOL[precision]
        assert perr <= PRE_REL[precision] * scale


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
@pytest.mark.parametrize("name,kw,th,tw", [("cfg3", K34, 61, 61), ("cfg4", K417, 129, 129), ("cfg5band", K241, 16, 129)])
def test_large_grid_windows_vs_oracle(name, kw, th, tw, precision):
    import infinite_texture_gans_b200 as itg
    ocfg = O.GenCfg(**kw)
    sd = O.make_state_dict(ocfg, seed=301, stress=True)
    z, maps = O.make_noise(ocfg, th, tw, seed=302)
    net = make_generator(kw, sd, precision, "cuda")
    img = itg.utils.generate_full_grid(net, z, maps).cpu()
    net.engine().drop_plans()
    torch.cuda.empty_cache()
    worst = 0.0
    for label, win in OW.standard_windows(th, tw).items():
        err, trusted = OW.window_error(sd, ocfg, z, maps, img, th, tw, win)
        worst = max(worst, err)
        print(f"{name} {precision} {label} {win} -> trusted {trusted}: max-abs {err:.3e}")
    assert worst <= TOL[precision], f"{name} {precision}: {worst:.3e}"


def test_fat_layers_with_many_work_items_match_direct_conv():
    """Streaming tcgen05 kernel on layers with K > 64 per tap and more than 2 x 148 work items (the configuration cfg3 / cfg4 /
    cfg5band run it in: two epilogue teams alternating accumulator buffers), against the emulator."""
    import math
    from emulator import EmulatorBackend
    from infinite_texture_gans_b200 import _lib as L, packing as PK
    from infinite_texture_gans_b200.ops import ConvOp, CudaBackend, Grid, c_store
    be = CudaBackend()
    dtype = torch.float16
    for mode, H, W, cin, cout in (("3x3", 64, 160, 208, 208), ("up", 40, 120, 416, 208), ("3x3", 96, 200, 104, 104)):
        g = torch.Generator().manual_seed(H + cin)
        kin, kout = c_store(cin), c_store(cout)
        src = Grid((torch.randn((H + 2, W + 2, kin), generator=g)).to(dtype), H, W, kin)
        wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
        w, s = (PK.pack_conv3x3(wt, dtype), 1) if mode == "3x3" else (PK.pack_upconv(wt, dtype), 2)
        oh, ow = s * H, s * W
        mk = lambda dev: ConvOp(mode=L.CONV3X3 if mode == "3x3" else L.UPCONV, src=Grid(src.buf.clone().to(dev), H, W, kin), w=w.to(dev), k=kin,
                                bias=PK.pad_vec(0.1 * torch.randn(cout, generator=torch.Generator().manual_seed(5)), w.shape[1]).to(dev),
                                impl=L.IMPL_UMMA, out_h=oh, out_w=ow, out_c=kout,
                                out_raw=Grid(torch.zeros((oh + 2, ow + 2, kout), dtype=dtype, device=dev), oh, ow, kout),
                                out_act=Grid(torch.zeros((oh + 2, ow + 2, kout), dtype=dtype, device=dev), oh, ow, kout),
                                scale=torch.ones(w.shape[1], device=dev), shift=torch.zeros(w.shape[1], device=dev), leak=0.02,
                                border=L.BORDER_REPLICATE, name=f"{mode}_{cin}")
        opc, opg = mk("cpu"), mk("cuda")
        EmulatorBackend().conv(opc)
        be.conv(opg)
        torch.cuda.synchronize()
        for a, b in ((opg.out_raw.interior, opc.out_raw.interior), (opg.out_act.buf, opc.out_act.buf)):
            ref = b.float()
            err = (a.float().cpu() - ref).abs().max().item()
            assert err <= 2.0 ** -9 * max(ref.abs().max().item(), 1e-6) + 1e-6, (mode, cin, err)
