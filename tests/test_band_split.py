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
