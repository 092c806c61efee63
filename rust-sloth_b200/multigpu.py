"""One huge frame across the GPUs of a box: destination-row bands, assembled without a gather step.

Every rank keeps the whole scene resident and owns rows [e[r], e[r+1]) (``sloth_ctx_set_band``): its geometry
kernel skips the chunks of 32 triangles whose bounding sphere cannot reach those rows and keeps only fragments
that land in the band.  Two ways to put the bands together:

``peer`` (default when the band offsets are 16-byte aligned)
    The root rank allocates the frame (two of them, used alternately) and exports CUDA IPC handles; every rank
    maps them and renders its band straight to ``frame + 4*e[r]*W``: the resolve kernel's stores travel over
    NVLink / NVSwitch while it runs, so there is no collective on the data path at all.  A 4-byte all-reduce per
    frame (NCCL, on the stream) tells the root that every band has landed and keeps ranks from running more than
    one frame ahead of the root.

``allgather``
    Each rank resolves into a local buffer and one ``all_gather_into_tensor`` moves 4*W*H/N bytes per GPU.

``host`` (:class:`HostBandRenderer`)
    The frame's destination is the host anyway (``Context::flush`` reads it there), so no GPU collects it: the
    ranks map one POSIX shared-memory frame, page-lock it (``sloth_host_register``) and every rank's
    ``sloth_render`` copies its band over its own PCIe link to ``frame + e[r]*W``.  Completion is a per-rank
    frame counter in the same shared segment (no collective, no NVLink traffic).  Band edges can follow the work:
    :meth:`HostBandRenderer.rebalance` moves them to equal shares of the chunks each band processed.

torch is only used for the process group, streams and (allgather mode) device memory.
"""
from __future__ import annotations

import numpy as np

from .turntable import band_edges


class BandRenderer:
    def __init__(self, ctx, width: int, height: int, rank: int, world: int, group=None, mode: str | None = None,
                 root: int = 0, depth: int = 1):
        import torch
        import torch.distributed as dist
        from . import device_alloc, ipc_export, ipc_open
        self.torch, self.dist, self.group = torch, dist, group
        self.ctx, self.W, self.H, self.rank, self.world, self.root = ctx, width, height, rank, world, root
        self.edges = band_edges(height, world)
        self.rows_max = max(self.edges[i + 1] - self.edges[i] for i in range(world))
        ctx.resize(width, height)
        ctx.set_band(self.edges[rank], self.edges[rank + 1])
        dev = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
        aligned = all((e * width) % 4 == 0 for e in self.edges)
        if mode is None:
            mode = "peer" if (world > 1 and aligned) else "allgather"
        if mode == "peer" and not aligned:
            raise ValueError("peer mode needs band offsets that are multiples of 4 cells (16 bytes)")
        self.mode = mode
        self.frame_cells = width * height + height          # image-mode layout: H trailing blank cells
        self.stride = (self.frame_cells + 3) & ~3           # frame slots stay 16-byte aligned
        self.depth = max(1, int(depth))                     # frames per render_batch call
        self.frames, self._owned, self._mapped = [], [], []
        self.k = 0
        if mode == "peer":
            # two regions of `depth` frame slots, used alternately: while the root reads one region the ranks
            # may already fill the other one
            n_slots = 2 * self.depth
            handles = [None]
            if rank == root:
                base = device_alloc(ctx.device, 4 * self.stride * n_slots)
                blank = np.full(self.stride, ord(" "), np.uint32)
                for i in range(n_slots):
                    ctx.write_device(base + 4 * self.stride * i, blank)
                self._owned.append(base)
                handles[0] = ipc_export(ctx.device, base)
            if world > 1:
                dist.broadcast_object_list(handles, src=root, group=group)
            if rank != root:
                base = ipc_open(ctx.device, handles[0])
                self._mapped.append(base)
            self.frames = [base + 4 * self.stride * i for i in range(n_slots)]
            self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            # equal-sized slots so that one all_gather_into_tensor moves everything
            self.local = torch.full((self.rows_max * width,), ord(" "), dtype=torch.int32, device=dev)
            self.all = torch.empty((world * self.rows_max * width,), dtype=torch.int32, device=dev)

    def close(self):
        from . import device_free, ipc_close
        self.torch.cuda.synchronize()
        if self.world > 1 and self.mode == "peer":
            self.dist.barrier(group=self.group)             # nobody is still writing into the root's frames
        for p in self._mapped:
            ipc_close(self.ctx.device, p)
        self._mapped = []
        if self.world > 1 and self.mode == "peer":
            self.dist.barrier(group=self.group)             # all mappings are closed before the owner frees
        for p in self._owned:
            device_free(self.ctx.device, p)
        self._owned = []

    def render(self, rot: np.ndarray):
        """Enqueue one frame.  peer mode: returns the device pointer of the assembled frame (meaningful on the
        root; complete once torch's current stream has passed this call).  allgather mode: returns the gathered
        device tensor (world * rows_max * W int32; rows of band i start at i*rows_max*W)."""
        torch, dist = self.torch, self.dist
        cur = torch.cuda.current_stream()
        if self.mode == "peer":
            frame = self.frames[(self.k & 1) * self.depth]
            self.k += 1
            self.stream.wait_stream(cur)                    # the previous frame's completion signal
            self.ctx.render_device(rot, frame + 4 * self.edges[self.rank] * self.W)
            cur.wait_stream(self.stream)
            if self.world > 1:
                dist.all_reduce(self.flag, group=self.group)   # every band of this frame has landed
            return frame
        self.ctx.render_device(rot, self.local.data_ptr())
        cur.wait_stream(self.stream)
        if self.world > 1:
            dist.all_gather_into_tensor(self.all, self.local, group=self.group)
        else:
            self.all.copy_(self.local)
        return self.all

    def render_batch(self, rots: np.ndarray):
        """peer mode only: up to `depth` frames in one call.  Inside it the geometry of frame k+1 overlaps the
        resolve of frame k, i.e. the NVLink stores of one frame are hidden behind the next frame's geometry, and
        there is one completion signal for the whole batch.  Returns the frame pointers (root)."""
        if self.mode != "peer":
            raise ValueError("render_batch needs peer mode")
        rots = np.ascontiguousarray(rots, np.float32).reshape(-1, 16)
        n = rots.shape[0]
        if n > self.depth:
            raise ValueError(f"{n} frames in one batch, but the renderer was created with depth={self.depth}")
        torch, dist = self.torch, self.dist
        cur = torch.cuda.current_stream()
        first = (self.k & 1) * self.depth
        self.k += 1
        self.stream.wait_stream(cur)
        self.ctx.render_device_batch(rots, self.frames[first] + 4 * self.edges[self.rank] * self.W, self.stride)
        cur.wait_stream(self.stream)
        if self.world > 1:
            dist.all_reduce(self.flag, group=self.group)
        return self.frames[first:first + n]

    def to_frame(self, result, image: bool = True) -> np.ndarray:
        """Host cell buffer in the reference's layout (W*H cells, +H blank cells in image mode); call it on the
        root in peer mode."""
        if self.mode == "peer":
            self.torch.cuda.current_stream().synchronize()
            return self.ctx.read_device(result, self.frame_cells if image else self.W * self.H)
        g = result.cpu().numpy().view(np.uint32)
        parts = [g[i * self.rows_max * self.W:i * self.rows_max * self.W + (self.edges[i + 1] - self.edges[i]) * self.W]
                 for i in range(self.world)]
        cells = np.concatenate(parts)
        if image:
            cells = np.concatenate([cells, np.full(self.H, ord(" "), np.uint32)])
        return cells


class HostBandRenderer:
    """Row bands assembled in a page-locked shared-memory host frame; one process per GPU.

    ``name`` identifies the shared segment (rank 0 creates it).  ``render(rot)`` returns when this rank's band is
    in the frame; ``wait(k)`` (any rank, usually the consumer on rank 0) returns when all bands of frame k are."""

    HEADER = 4096   # bytes: per-rank frame counters (uint64), then the frame

    def __init__(self, ctx, width: int, height: int, rank: int, world: int, name: str, barrier=None, edges=None,
                 page_lock: bool = True):
        from multiprocessing import shared_memory
        from . import host_register
        self.page_lock = page_lock
        self.ctx, self.W, self.H, self.rank, self.world = ctx, width, height, rank, world
        self.barrier = barrier or (lambda: None)
        self.frame_cells = width * height + height
        nbytes = self.HEADER + 4 * self.frame_cells
        if rank == 0:
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=nbytes)
            self.shm.buf[:self.HEADER] = bytes(self.HEADER)
        self.barrier()
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
        self.counters = np.ndarray((world,), np.uint64, self.shm.buf, 0)
        self.frame = np.ndarray((self.frame_cells,), np.uint32, self.shm.buf, self.HEADER)
        self._whole = np.ndarray((nbytes,), np.uint8, self.shm.buf, 0)
        if page_lock:
            host_register(self._whole)
        if rank == 0:
            self.frame[width * height:] = ord(" ")          # image-mode tail, context.rs:38-39
        ctx.resize(width, height)
        self.k = 0
        self.set_edges(edges or band_edges(height, world))

    def set_edges(self, edges):
        self.edges = [int(e) for e in edges]
        assert self.edges[0] == 0 and self.edges[-1] == self.H and all(a < b for a, b in zip(self.edges, self.edges[1:]))
        self.ctx.set_band(self.edges[self.rank], self.edges[self.rank + 1])
        self.barrier()

    def render(self, rot: np.ndarray) -> int:
        r0 = self.edges[self.rank]
        self.ctx.render_into(rot, self.frame[r0 * self.W:])
        self.k += 1
        self.counters[self.rank] = self.k                   # the copy has completed (sloth_render synchronises)
        return self.k

    def wait(self, k: int) -> np.ndarray:
        while int(self.counters.min()) < k:
            pass
        return self.frame

    def rebalance(self, chunks_per_band) -> list[int]:
        """New edges with equal shares of the work, from the chunks each band processed under the current edges
        (work taken as uniform inside a band).  All ranks must call it with the same numbers."""
        w = np.maximum(np.asarray(chunks_per_band, np.float64), 1.0)
        cdf_rows = np.array(self.edges, np.float64)
        cdf_work = np.concatenate([[0.0], np.cumsum(w)])
        targets = cdf_work[-1] * np.arange(1, self.world) / self.world
        cuts = np.interp(targets, cdf_work, cdf_rows)
        edges = [0] + [int(round(c)) for c in cuts] + [self.H]
        for i in range(1, len(edges)):                      # strictly increasing
            edges[i] = max(edges[i], edges[i - 1] + 1)
        edges[-1] = self.H
        for i in range(len(edges) - 2, 0, -1):
            edges[i] = min(edges[i], edges[i + 1] - 1)
        self.set_edges(edges)
        return edges

    def close(self):
        from . import host_unregister
        self.barrier()
        if self.page_lock:
            host_unregister(self._whole)
        del self.counters, self.frame, self._whole
        self.shm.close()
        self.barrier()
        if self.rank == 0:
            self.shm.unlink()
