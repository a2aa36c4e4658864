"""Frame patching at the halo points of a plan (engine.py).

Two users:

* `SequentialHalo` -- the eval-mode state machine of `LocalPadder` (models/layers.py:78-143): when a large
  texture is produced as a sequence of nph x npw sub-images, every conv2d_lp input keeps the pixel column
  that the next sub-image in the row needs as its left halo and the pixel row that the next row of
  sub-images needs as its top halo.  The reference parks the rows on the host; here they stay on the device
  and are written straight into the frame of the consumer's grid tensor.
* `BandHalo` -- the row-band multi-GPU split: the top / bottom frame rows of every conv2d_lp input are the
  neighbour ranks' border rows, exchanged with NCCL send/recv (or gloo on CPU for the tests).

Both run right after the producer of the grid tensor (which has already written the outer padding into
the whole frame) and before its consumer conv.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .engine import HaloPoint, Plan


@dataclass
class _LayerState:
    """State of ONE conv2d_lp input (models/layers.py:69-76), all tensors on the device."""
    col: Optional[torch.Tensor] = None         # (H+2, 1, C): left halo of the current sub-image, corners included
    col_next: Optional[torch.Tensor] = None
    row_parts: List[torch.Tensor] = field(default_factory=list)   # slices collected for the next row of sub-images
    row_cur: Optional[torch.Tensor] = None     # (1, total_w*w + 2, C): top halo of the current row, padded by 1 px
    row_off: int = 0


class SequentialHalo:
    """image_location protocol (utils.py:321-337; substring tests of models/layers.py:81-140)."""

    def __init__(self, backend, nph: int, npw: int, replicate: bool):
        self.be, self.nph, self.npw, self.replicate = backend, nph, npw, replicate
        self.layers: Dict[str, _LayerState] = {}
        self._pool: Dict[tuple, List[torch.Tensor]] = {}      # retired halo buffers by (shape, dtype, device): steady state allocates nothing

    def reset(self) -> None:
        self.layers.clear()

    def _get(self, shape, like: torch.Tensor) -> torch.Tensor:
        free = self._pool.get((tuple(shape), like.dtype, like.device))
        return free.pop() if free else torch.empty(shape, dtype=like.dtype, device=like.device)

    def _put(self, t: Optional[torch.Tensor]) -> None:
        if t is not None:
            self._pool.setdefault((tuple(t.shape), t.dtype, t.device), []).append(t)

    def hooks(self, plan: Plan, loc: str):
        return {hp.step: (lambda hp=hp: self.apply(hp, loc)) for hp in plan.halo_points}

    def apply(self, hp: HaloPoint, loc: str) -> None:
        be, g, w = self.be, hp.grid, hp.r
        buf, H, W, C = g.buf, g.h, g.w, g.c
        st = self.layers.setdefault(hp.name, _LayerState())
        first_row, first_col, last_col = "1st_row" in loc, "1st_col" in loc, "last_col" in loc
        if not (first_row and first_col) and st.col is None and st.col_next is None and st.row_cur is None \
                and not st.row_parts:
            # models/layers.py:86 fails with a TypeError on the None halo; say what is wrong instead
            raise RuntimeError(f"image_location {loc!r} needs halos from earlier sub-images, but none are stored "
                               f"(start a sweep with '1st_row_1st_col')")

        # ---- update_padding_variables (models/layers.py:103-143) ----
        if st.col_next is not None:
            self._put(st.col)
            st.col, st.col_next = st.col_next, None
        if not last_col:
            # column W*(npw-1)-1 of the merged input; frame rows travel along and become the corners
            nxt = self._get((H + 2, 1, C), buf)
            be.copy_rect(buf, 0, w * (self.npw - 1), nxt, 0, 0, H + 2, 1)
            st.col_next = nxt
        ncols = W if last_col else w * (self.npw - 1)
        part = self._get((1, ncols, C), buf)
        be.copy_rect(buf, w * (self.nph - 1), 1, part, 0, 0, 1, ncols)      # row H*(nph-1)-1
        if first_col:
            if not first_row:
                total = sum(p.shape[1] for p in st.row_parts)
                cur = self._get((1, total + 2, C), buf)
                x = 1
                for p in st.row_parts:
                    be.copy_rect(p, 0, 0, cur, 0, x, 1, p.shape[1])
                    x += p.shape[1]
                    self._put(p)
                if self.replicate:                                           # F.pad(row, (1,1,0,0), outer_padding)
                    be.copy_rect(cur, 0, 1, cur, 0, 0, 1, 1)
                    be.copy_rect(cur, 0, total, cur, 0, total + 1, 1, 1)
                else:
                    cur[:, 0].zero_()
                    cur[:, total + 1].zero_()
                st.row_cur, st.row_off = cur, 0
            st.row_parts = [part]
        else:
            st.row_parts.append(part)

        # ---- padding (models/layers.py:78-101): outer padding is already in the frame ----
        if first_row and first_col:
            pass
        else:
            if not first_col:
                if st.col is None:
                    raise RuntimeError(f"{hp.name}: no stored left halo for image_location {loc!r}")
                be.copy_rect(st.col, 0, 0, buf, 0, 0, H + 2, 1)
        if st.row_cur is not None:
            if not first_row:
                be.copy_rect(st.row_cur, 0, st.row_off, buf, 0, 0, 1, W + 2)
            if last_col:
                self._put(st.row_cur)
                st.row_cur = None
            else:
                st.row_off += (self.npw - 1) * w


class BandHalo:
    """Row-band split of one texture over the ranks of a process group: rank i owns patch rows
    [row0_i, row1_i); every conv2d_lp input sends its first / last interior pixel row (frame columns
    included) to the rank above / below and receives their rows into its own top / bottom frame row."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bytes_sent = 0

    def hooks(self, plan: Plan):
        if self.world == 1:
            return None
        return {hp.step: (lambda hp=hp: self.exchange(hp)) for hp in plan.halo_points}

    def exchange(self, hp: HaloPoint) -> None:
        dist, buf, H = self.dist, hp.grid.buf, hp.grid.h
        ops = []
        up = self.rank - 1 if self.rank > 0 else None
        down = self.rank + 1 if self.rank < self.world - 1 else None
        gr = (lambda r: dist.get_global_rank(self.group, r)) if self.group is not None else (lambda r: r)
        if up is not None:
            ops.append(dist.P2POp(dist.isend, buf[1], gr(up), self.group))
            ops.append(dist.P2POp(dist.irecv, buf[0], gr(up), self.group))
        if down is not None:
            ops.append(dist.P2POp(dist.isend, buf[H], gr(down), self.group))
            ops.append(dist.P2POp(dist.irecv, buf[H + 1], gr(down), self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        self.bytes_sent += buf[1].numel() * buf.element_size() * ((up is not None) + (down is not None))


class P2PBandHalo:
    """Row-band split with the halo rows moved by the GPUs themselves over peer-mapped memory (NVLink P2P):
    one `itg_halo_exchange` launch per conv2d_lp input pushes this rank's border rows into the neighbours' inboxes,
    publishes a step number in their flags, waits for the neighbours' flags and pulls the received rows into the
    frame.  No host synchronisation and no NCCL call inside a step, so a whole step (launches + exchanges) can be
    captured in one CUDA graph.  torch.distributed is only used once, to swap the CUDA IPC handles."""

    FLAG_BYTES = 4096

    def __init__(self, plan: Plan, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib as L
        self.lib = L.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.plan = plan
        dev = plan.out.device
        self.device_index = dev.index if dev.index is not None else torch.cuda.current_device()
        es = plan.halo_points[0].grid.buf.element_size()
        self.dtype_code = L.DTYPE_OF[plan.halo_points[0].grid.buf.dtype]
        # layout of the exchange buffer: [flags + step counter | per halo point: top inbox row, bottom inbox row]
        self.offsets = []
        off = self.FLAG_BYTES
        for hp in plan.halo_points:
            row = ((hp.grid.w + 2) * hp.grid.c * es + 255) // 256 * 256
            self.offsets.append((off, off + row))
            off += 2 * row
        self.bytes = off
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        L.check(self.lib.itg_ipc_alloc(self.device_index, self.bytes, C.byref(ptr), handle))
        self.base = ptr.value
        infos = [None] * self.world
        dist.all_gather_object(infos, (self.device_index, handle.raw, self.bytes), group)
        if any(i[2] != self.bytes for i in infos):
            raise RuntimeError("row bands must have the same shape on every rank for the P2P exchange")
        self.peer = {}
        for nb in (self.rank - 1, self.rank + 1):
            if 0 <= nb < self.world:
                p = C.c_void_p()
                L.check(self.lib.itg_ipc_open(self.device_index, infos[nb][1], C.byref(p)))
                self.peer[nb] = p.value
        dist.barrier(group)
        self.step_ptr = self.base + 8 * len(plan.halo_points) + 64      # int32 step counter after the flags
        self.bytes_per_step = sum((hp.grid.w + 2) * hp.grid.c * es for hp in plan.halo_points) * len(self.peer)   # pushed by this rank

    def _args(self, k: int, hp: HaloPoint):
        up, down = self.peer.get(self.rank - 1), self.peer.get(self.rank + 1)
        top_off, bot_off = self.offsets[k]
        flag = lambda base, i: base + 4 * i
        return (self.dtype_code, hp.grid.buf.data_ptr(), hp.grid.h, hp.grid.w, hp.grid.c,
                (up + bot_off) if up else None, (down + top_off) if down else None,          # my first row -> up neighbour's BOTTOM inbox
                flag(up, 2 * k + 1) if up else None, flag(down, 2 * k) if down else None,
                (self.base + top_off) if up else None, (self.base + bot_off) if down else None,
                flag(self.base, 2 * k) if up else None, flag(self.base, 2 * k + 1) if down else None,
                self.step_ptr)

    def hooks(self, plan: Plan):
        from . import _lib as L
        assert plan is self.plan
        if self.world == 1:
            return None
        lib = self.lib
        per_step = {}
        for k, hp in enumerate(plan.halo_points):
            args = self._args(k, hp)
            if hp.pull_step == hp.step:          # consumer follows immediately: one launch pushes and pulls
                per_step.setdefault(hp.step, []).append((args, 15))
            else:                                # independent work in between: push now, pull right before the consumer
                per_step.setdefault(hp.step, []).append((args, 3))
                per_step.setdefault(hp.pull_step, []).append((args, 12))

        def make(calls):
            def run():
                for args, roles in calls:
                    L.check(lib.itg_halo_exchange(*args, roles, L.stream_ptr()))
            return run
        return {step: make(calls) for step, calls in per_step.items()}

    def begin_step(self) -> None:
        """Advance the device-resident step counter (first launch of every Generator pass)."""
        from . import _lib as L
        L.check(self.lib.itg_step_advance(self.step_ptr, L.stream_ptr()))

    def close(self) -> None:
        for p in self.peer.values():
            self.lib.itg_ipc_close(p)
        self.peer = {}
        if self.base:
            self.lib.itg_ipc_free(self.base)
            self.base = 0
