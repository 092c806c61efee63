"""One huge frame across the GPUs of a box: destination-row bands + one NCCL all-gather.

Every rank keeps the whole scene resident, owns rows [e[r], e[r+1]) (``sloth_ctx_set_band``), runs
geometry over all triangles but keeps only fragments that land in its band, resolves its band into a
device buffer, and the bands are exchanged with ``all_gather_into_tensor`` over NVLink (4*W*H/N bytes
per GPU).  torch is only used for device memory and the collective.
"""
from __future__ import annotations

import numpy as np

from .turntable import band_edges


class BandRenderer:
    def __init__(self, ctx, width: int, height: int, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.ctx, self.W, self.H, self.rank, self.world = ctx, width, height, rank, world
        self.edges = band_edges(height, world)
        self.rows_max = max(self.edges[i + 1] - self.edges[i] for i in range(world))
        ctx.resize(width, height)
        ctx.set_band(self.edges[rank], self.edges[rank + 1])
        dev = torch.device("cuda", ctx.device)
        # equal-sized slots so that one all_gather_into_tensor moves everything
        self.local = torch.full((self.rows_max * width,), ord(" "), dtype=torch.int32, device=dev)
        self.all = torch.empty((world * self.rows_max * width,), dtype=torch.int32, device=dev)
        self.stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)

    def render(self, rot: np.ndarray):
        """Returns the gathered device tensor (world * rows_max * W int32); rows of band i start at
        i*rows_max*W."""
        torch, dist = self.torch, self.dist
        self.ctx.render_device(rot, self.local.data_ptr())
        ev = torch.cuda.Event()
        ev.record(self.stream)
        torch.cuda.current_stream().wait_event(ev)
        if self.world > 1:
            dist.all_gather_into_tensor(self.all, self.local, group=self.group)
        else:
            self.all.copy_(self.local)
        return self.all

    def to_frame(self, gathered, image: bool = True) -> np.ndarray:
        """Host cell buffer in the reference's layout (W*H cells, +H blank cells in image mode)."""
        g = gathered.cpu().numpy().view(np.uint32)
        parts = [g[i * self.rows_max * self.W:i * self.rows_max * self.W + (self.edges[i + 1] - self.edges[i]) * self.W]
                 for i in range(self.world)]
        cells = np.concatenate(parts)
        if image:
            cells = np.concatenate([cells, np.full(self.H, ord(" "), np.uint32)])
        return cells
