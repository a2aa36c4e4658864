"""Multi-GPU samplers, one process per GPU.

* `RowBandSampler` / `sample_row_bands`: ONE seamless texture whose patch grid is split into row bands of whole patch rows, one band per
  rank.  Every op of the Generator is pointwise, a 3x3 stencil or per-patch attention, so the only exchange is one pixel row per
  conv2d_lp input and direction (halo.P2PBandHalo: peer-mapped memory over NVLink, no host involvement, the whole step in one CUDA
  graph; halo.BandHalo: torch.distributed send/recv -- NCCL, or gloo for the CPU tests).  This replaces the reference's way of
  producing a large image -- `utils.sample_from_gen_PatchByPatch_test` (utils.py:258-397) walking 3x3 sub-images on one device with the
  halo rows parked on the host (models/layers.py:117-139).
* `generate_textures_replicas`: many independent textures (BASELINE.json config 4), rank r produces textures r, r + world, ...: replicas
  only, no collective on the data path.

Noise: every rank needs only its band of the full-grid noise (`band_noise` slices host tensors drawn in the reference's order,
utils.py:228, 246; `utils.draw_noise_device` generates exactly the band on the device from a counter-based generator).
"""
from __future__ import annotations

import itertools
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .config import GenConfig


def split_rows(total_rows: int, world: int) -> List[Tuple[int, int]]:
    """Patch rows [r0, r1) of each rank: bands of whole patch rows, as equal as possible (the first total_rows % world get one more)."""
    if world < 1 or total_rows < world:
        raise ValueError(f"cannot split {total_rows} patch rows over {world} ranks (every rank needs at least one row)")
    base, extra = divmod(total_rows, world)
    out, r0 = [], 0
    for r in range(world):
        h = base + (1 if r < extra else 0)
        out.append((r0, r0 + h))
        r0 += h
    return out


def band_noise(cfg: GenConfig, z_full: torch.Tensor, maps_full: Optional[Sequence[torch.Tensor]], r0: int, r1: int):
    """Rows [r0, r1) of the full-grid noise with the rings a band shares with its neighbours: z (1, z_dim, total_h*b+2, total_w*b+2)
    -> (z_dim, (r1-r0)*b+2, W); per level i, map (1, 1, total_h*r_i+4, W_i) -> ((r1-r0)*r_i+4, W_i).  (utils.py:228-256 geometry.)"""
    b = cfg.base_res
    z = z_full[0] if z_full.dim() == 4 else z_full
    zb = z[:, r0 * b:r1 * b + 2].contiguous()
    mb = None
    if maps_full is not None:
        mb = []
        for i, m in enumerate(maps_full):
            m2 = m[0, 0] if m.dim() == 4 else m
            mb.append(m2[r0 * b * 2 ** i:r1 * b * 2 ** i + 4].contiguous())
    return zb, mb


class RowBandSampler:
    """This rank's band of a total_h x total_w patch texture.

        s = RowBandSampler(netG, total_h, total_w)          # collective: every rank of `group` constructs it
        s.set_noise(z_full, maps_full)                      # or s.set_band_noise(z_band, maps_band)
        img = s.step()                                      # (1, img_ch, rows*P, total_w*P) device tensor: this rank's band
        s.close()

    halo: 'p2p' (device-side exchange over peer-mapped memory; needs CUDA IPC between the ranks' GPUs), 'dist' (torch.distributed
    send/recv per conv input; also the CPU / gloo path) or 'auto' (p2p when every rank can set it up, else dist).
    graph: capture launches + exchanges of one step into a CUDA graph (p2p only: a NCCL exchange inside a capture hung in round 1)."""

    def __init__(self, netG, total_h: int, total_w: int, group=None, halo: str = "auto", graph: bool = True):
        import torch.distributed as dist
        from .halo import BandHalo, P2PBandHalo
        from .utils import _unwrap
        if halo not in ("auto", "p2p", "dist"):
            raise ValueError("halo must be 'auto', 'p2p' or 'dist'")
        self.dist, self.group = dist, group
        self.G = _unwrap(netG)
        self.cfg: GenConfig = self.G.cfg
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.total_h, self.total_w = total_h, total_w
        self.rows = split_rows(total_h, self.world)[self.rank]
        th = self.rows[1] - self.rows[0]
        self.eng = self.G.engine()
        self.plan = self.eng.plan(th, total_w, L.IMG_MERGED)
        self.device = self.plan.out.device
        self.p2p, self.band = False, None
        if self.world > 1:
            if halo in ("auto", "p2p") and self.device.type == "cuda":
                err = None
                try:
                    self.band, self.p2p = P2PBandHalo(self.plan, group), True
                except Exception as e:                                   # noqa: BLE001  (no IPC between these devices)
                    err = e
                ok = torch.tensor([1 if self.p2p else 0], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 0:
                    if self.p2p:
                        self.band.close()
                    self.band, self.p2p = None, False
                    if halo == "p2p":
                        raise RuntimeError(f"P2P halo exchange could not be set up on every rank ({err})")
            if self.band is None:
                self.band = BandHalo(group)
        self.hooks = self.band.hooks(self.plan) if self.band is not None else None
        self.use_graph = bool(graph) and self.device.type == "cuda" and (self.p2p or self.world == 1)
        self._graph = None
        # launches of one step (the library's own kernels): the plan's, one exchange per halo point and direction pair, the step counter
        self.launches_per_step = self.plan.n_launches + ((self._n_exchange_launches() + 1) if self.p2p else 0)

    def _n_exchange_launches(self) -> int:
        return sum(1 if hp.pull_step == hp.step else 2 for hp in self.plan.halo_points)

    # ---- inputs ----
    def set_band_noise(self, z_band: torch.Tensor, maps_band: Optional[Sequence[torch.Tensor]] = None) -> None:
        with self.eng._on_device():
            self.plan.set_inputs(z_band, maps_band)

    def set_noise(self, z_full: torch.Tensor, maps_full: Optional[Sequence[torch.Tensor]] = None) -> None:
        zb, mb = band_noise(self.cfg, z_full, maps_full, *self.rows)
        self.set_band_noise(zb, mb)

    # ---- one Generator pass over the band ----
    def _eager(self) -> None:
        if self.p2p:
            self.band.begin_step()
        self.plan.run(self.hooks)

    def step(self) -> torch.Tensor:
        with self.eng._on_device():
            if not self.use_graph:
                self._eager()
            else:
                if self._graph is None:
                    self._eager()                                        # warm-up outside capture (one-time attribute / tensor-map setup)
                    torch.cuda.current_stream().synchronize()
                    if self.world > 1:
                        self.dist.barrier(self.group)                    # ranks aligned before the first captured exchange
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._eager()
                    self._graph = g
                self._graph.replay()
        return self.plan.out

    def generate(self, band_noises: Iterable, out_format: str = "float32") -> Iterator[torch.Tensor]:
        """Streaming counterpart of `utils.generate_textures` for this rank's band: one host-resident band image per element of
        `band_noises` ((z_band, maps_band) per step, ideally pinned).  Step k's image crosses PCIe while step k+1 computes and step
        k+2's noise is uploaded; every rank must consume the same number of steps (the halo exchange pairs them up).
        out_format: 'float32' -> (1, img_ch, rows*P, W*P) fp32; 'uint8' -> (rows*P, W*P, img_ch) bytes of test_sample.py's output stage."""
        from . import utils as U
        if out_format not in ("float32", "uint8"):
            raise ValueError("out_format must be 'float32' or 'uint8'")
        if self.device.type != "cuda":
            raise L.ItgError("RowBandSampler.generate stages through pinned host memory: CUDA devices only")
        shape = tuple(self.plan.out.shape)
        pipes = self.__dict__.setdefault("_pipes", {})
        pipe = pipes.get(out_format)
        if pipe is None:
            pipe = pipes[out_format] = (U.HostOutputPipe((shape[2], shape[3], shape[1]), self.device, dtype=torch.uint8) if out_format == "uint8"
                                        else U.HostOutputPipe(shape, self.device))
        up = self.__dict__.get("_uploader")
        if up is None:
            up = self.__dict__["_uploader"] = U.NoiseUploader(self.plan, self.device)
        up.head = up.tail = 0
        it = iter(band_noises)
        nxt = next(it, None)
        if nxt is None:
            return
        with torch.cuda.device(self.device):
            up.upload(nxt)
        in_flight: List[int] = []
        while nxt is not None:
            with torch.cuda.device(self.device):
                up.feed()
                img = self.step()
                nxt = next(it, None)
                if nxt is not None:
                    up.upload(nxt)
                in_flight.append(pipe.push(img))
            if len(in_flight) == pipe.depth:
                yield pipe.wait(in_flight.pop(0))
        for slot in in_flight:
            yield pipe.wait(slot)

    def gather(self, band_img: Optional[torch.Tensor] = None, dst: int = 0) -> Optional[torch.Tensor]:
        """Assemble the full (1, img_ch, total_h*P, total_w*P) image on rank `dst` (collective; bands may differ in height)."""
        dist = self.dist
        img = (self.plan.out if band_img is None else band_img).contiguous()
        if self.world == 1:
            return img
        P = self.cfg.patch_px
        if self.rank == dst:
            parts = []
            for r, (a, b) in enumerate(split_rows(self.total_h, self.world)):
                if r == dst:
                    parts.append(img)
                else:
                    buf = torch.empty((1, self.cfg.img_ch, (b - a) * P, self.total_w * P), dtype=img.dtype, device=img.device)
                    dist.recv(buf, src=r if self.group is None else dist.get_global_rank(self.group, r), group=self.group)
                    parts.append(buf)
            return torch.cat(parts, dim=2)
        dist.send(img, dst=dst if self.group is None else dist.get_global_rank(self.group, dst), group=self.group)
        return None

    def close(self) -> None:
        self._graph = None
        if self.p2p and self.band is not None:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            self.dist.barrier(self.group)                                # nobody unmaps a buffer a neighbour may still write
            self.band.close()
        self.band = None


def sample_row_bands(netG, output_resolution_height: int, output_resolution_width: int, noise=None, group=None, halo: str = "auto",
                     gather: bool = True, seed: Optional[int] = None) -> Optional[torch.Tensor]:
    """Multi-GPU counterpart of `utils.sample_from_gen_PatchByPatch_test` (utils.py:258-397) for one large texture: every rank of
    `group` calls it; the patch grid (same geometry as the reference sampler) is split into row bands.  noise: the full-grid
    (z_full, maps_full) on the host (identical on every rank), or None to let every rank generate exactly its own band on the device
    from the counter-based generator (`seed`, identical on every rank).  Returns, on rank 0 (gather=True), the (1, img_ch, H, W) image on
    the device, None elsewhere; with gather=False every rank gets its own band (uncropped)."""
    from . import utils as U
    G = U._unwrap(netG)
    geo = U.patch_grid_geometry(output_resolution_height, output_resolution_width, G.n_layers_G, G.cfg.base_res)
    th, tw = geo["total_h"], geo["total_w"]
    s = RowBandSampler(netG, th, tw, group=group, halo=halo, graph=False)
    try:
        if noise is not None:
            s.set_noise(*noise)
        else:
            if seed is None:
                raise ValueError("pass either the full-grid noise or a seed for the device-side generator")
            zb, mb = U.draw_noise_device(G.cfg, th, tw, seed, rows=s.rows, device=s.device)
            s.set_band_noise(zb, mb)
        band = s.step()
        if not gather:
            return band.clone()
        full = s.gather(band)
        return None if full is None else full[:, :, :output_resolution_height, :output_resolution_width]
    finally:
        s.close()


def generate_textures_replicas(netG, noises: Iterable, output_resolution_height: int, output_resolution_width: int, group=None,
                               rank: Optional[int] = None, world: Optional[int] = None, **kw) -> Iterator[Tuple[int, torch.Tensor]]:
    """Independent textures sharded over the ranks (BASELINE.json config 4: 256 textures over 8 GPUs = 32 per GPU): rank r yields
    (index, image) for textures r, r + world, r + 2*world, ... of `noises` through `utils.generate_textures` (copies overlapped with
    compute).  Replicas only: the Generator's weights are replicated, no collective runs on the data path.  The reference has no
    batch or multi-device sampler at all (utils.py:341 ignores num_images)."""
    from . import utils as U
    if rank is None or world is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    mine = itertools.islice(noises, rank, None, world)
    for k, img in enumerate(U.generate_textures(netG, mine, output_resolution_height, output_resolution_width, **kw)):
        yield rank + k * world, img
