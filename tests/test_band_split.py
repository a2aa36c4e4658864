"""Row-band multi-GPU split (halo.BandHalo) on CPU: two gloo ranks, each running the launch plan of its own band of
patch rows on the launch emulator and exchanging one pixel row per conv2d_lp input with its neighbour.  The bands,
stacked, must equal the single-process result bit for bit (same launches, same reduction order) and match the
reference's golden output."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import compare_with_golden, load_case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, name, splits, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emulator import EmulatorBackend
        from infinite_texture_gans_b200 import _lib as L
        from infinite_texture_gans_b200.config import GenConfig
        from infinite_texture_gans_b200.engine import Engine
        from infinite_texture_gans_b200.halo import BandHalo
        d, kw, ocfg, sd, z, maps = load_case(name)
        cfg = GenConfig(**kw)
        tw = int(d["total_w"])
        r0, r1 = splits[rank], splits[rank + 1]
        th, b = r1 - r0, cfg.base_res
        eng = Engine(cfg, sd, "fp32", "cpu", backend=EmulatorBackend())
        plan = eng.plan(th, tw, L.IMG_MERGED)
        zb = z[0, :, r0 * b:r1 * b + 2].contiguous()
        mb = None
        if maps is not None:
            mb = [m[0, 0, r0 * b * 2 ** i:r1 * b * 2 ** i + 4].contiguous() for i, m in enumerate(maps)]
        plan.set_inputs(zb, mb)
        band = BandHalo()
        out = plan.run(band.hooks(plan)).clone()
        torch.save({"out": out, "bytes": band.bytes_sent}, os.path.join(out_dir, f"band{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,splits", [("gen_bn4_att_rep", (0, 3, 5)), ("gen_ssm4_att_rep", (0, 1, 3)),
                                         ("gen_bn4_noatt_const_crop", (0, 2, 3)),
                                         ("gen_bn4_att_rep", (0, 2, 3, 5))])           # three ranks: the middle one has two neighbours
def test_two_rank_band_split_equals_single_process(name, splits, tmp_path):
    world = len(splits) - 1
    mp.spawn(_rank_main, args=(world, _free_port(), name, splits, str(tmp_path)), nprocs=world, join=True)
    bands = [torch.load(os.path.join(str(tmp_path), f"band{r}.pt")) for r in range(world)]
    img = torch.cat([bd["out"] for bd in bands], dim=2)
    d, kw, ocfg, sd, z, maps = load_case(name)
    compare_with_golden(d, "one", img, 5e-5)
    # single-process plan of the whole grid: identical launches -> bit-identical pixels
    from emulator import EmulatorBackend
    from infinite_texture_gans_b200.config import GenConfig
    from infinite_texture_gans_b200.engine import Engine
    eng = Engine(GenConfig(**kw), sd, "fp32", "cpu", backend=EmulatorBackend())
    full = eng.forward(z, None if maps is None else [m[0, 0] for m in maps], th=int(d["total_h"]), tw=int(d["total_w"]))
    assert torch.equal(img, full)
    assert all(bd["bytes"] > 0 for bd in bands)


def _sampler_rank_main(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from common import make_generator
        from emulator import EmulatorBackend
        from infinite_texture_gans_b200 import bands
        d, kw, ocfg, sd, z, maps = load_case(name)
        net = make_generator(kw, sd, "fp32", backend=EmulatorBackend())
        s = bands.RowBandSampler(net, int(d["total_h"]), int(d["total_w"]))          # halo='auto' -> torch.distributed send/recv on CPU
        assert not s.p2p and s.rows == bands.split_rows(int(d["total_h"]), world)[rank]
        s.set_noise(z, maps)
        band = s.step()
        full = s.gather(band)
        assert (full is not None) == (rank == 0)
        if rank == 0:
            torch.save(full, os.path.join(out_dir, "full.pt"))
        s.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("gen_bn4_att_rep", 2), ("gen_ssm4_att_rep", 3)])
def test_row_band_sampler_public_api(name, world, tmp_path):
    """bands.RowBandSampler (the package's multi-GPU sampler for one large texture): unequal bands, gather on rank 0, result equal to the
    reference's golden one-shot image."""
    mp.spawn(_sampler_rank_main, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    d, kw, ocfg, sd, z, maps = load_case(name)
    compare_with_golden(d, "one", torch.load(os.path.join(str(tmp_path), "full.pt")), 5e-5)


def test_split_rows_and_band_noise():
    from infinite_texture_gans_b200 import bands
    from infinite_texture_gans_b200.config import GenConfig
    assert bands.split_rows(513, 8) == [(0, 65), (65, 129), (129, 193), (193, 257), (257, 321), (321, 385), (385, 449), (449, 513)]
    assert bands.split_rows(5, 2) == [(0, 3), (3, 5)] and bands.split_rows(4, 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    with pytest.raises(ValueError):
        bands.split_rows(3, 4)
    cfg = GenConfig(z_dim=8, G_ch=8, n_layers_G=4, type_norm="SSM")
    z = torch.arange(8 * 22 * 14, dtype=torch.float32).reshape(1, 8, 22, 14)              # 5 x 3 patches
    maps = [torch.arange((5 * r + 4) * (3 * r + 4), dtype=torch.float32).reshape(1, 1, 5 * r + 4, 3 * r + 4) for r in (4, 8, 16, 32)]
    zb, mb = bands.band_noise(cfg, z, maps, 2, 5)
    assert torch.equal(zb, z[0, :, 8:22]) and [tuple(m.shape) for m in mb] == [(16, 16), (28, 28), (52, 52), (100, 100)]
    assert torch.equal(mb[1], maps[1][0, 0, 16:44])
    # neighbouring bands share the ring rows (1 px of z, 2 px of every map on each side of the seam)
    za, ma = bands.band_noise(cfg, z, maps, 0, 2)
    assert torch.equal(za[:, -2:], zb[:, :2]) and torch.equal(ma[2][-4:], mb[2][:4])
