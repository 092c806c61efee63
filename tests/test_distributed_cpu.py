"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: frame sharding of a turntable
and row-band assembly of one frame.  The per-rank renderer here is the oracle (no GPU in this
container); on the GPU box tests/test_gpu_parity.py covers the same split with the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import scenes as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, W, H, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rust_sloth_b200  # noqa: F401
    from rust_sloth_b200 import turntable as tt
    xyz, rgb, s0 = S.soup("pikachu")
    rots = np.stack([oracle.rotation(0.0, p, 0.0) for p in oracle.turntable(0.0, n_frames)])
    # --- frames sharded round-robin, gathered to rank 0 in frame order -------------------------
    mine = tt.frame_shard(n_frames, rank, world)
    local = np.stack([oracle.render(xyz, rgb, s0, W, H, rots[k], mode=1)[0] for k in mine]).astype(np.int64)
    sizes = [len(tt.frame_shard(n_frames, r, world)) for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad, local.shape[1]), dtype=torch.int64)
    buf[:len(mine)] = torch.from_numpy(local)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    # --- one frame split in row bands, all-gathered ------------------------------------------------
    edges = tt.band_edges(H, world)
    full = oracle.render(xyz, rgb, s0, W, H, rots[1], mode=1)[0]
    band = torch.from_numpy(full[edges[rank] * W:edges[rank + 1] * W].astype(np.int64))
    rows = max(edges[i + 1] - edges[i] for i in range(world))
    bbuf = torch.zeros(rows * W, dtype=torch.int64)
    bbuf[:band.numel()] = band
    bands = [torch.zeros_like(bbuf) for _ in range(world)]
    dist.all_gather(bands, bbuf)
    parts = [bands[i][:(edges[i + 1] - edges[i]) * W].numpy().astype(np.uint32) for i in range(world)]
    whole = tt.assemble_bands(parts, W, H, image=True)
    ok_band = bool(np.array_equal(whole, full))
    if rank == 0:
        shards = [gathered[r][:sizes[r]].numpy().astype(np.uint32) for r in range(world)]
        frames = tt.interleave_shards(shards, n_frames, world)
        np.save(out_path, np.stack(frames))
    res = torch.tensor([1 if ok_band else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    assert int(res) == 1
    dist.destroy_process_group()


def test_frame_sharding_and_band_assembly_world2(tmp_path):
    n_frames, W, H = 7, 80, 41
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, W, H, out), nprocs=2, join=True)
    frames = np.load(out)
    xyz, rgb, s0 = S.soup("pikachu")
    for k, p in enumerate(oracle.turntable(0.0, n_frames)):
        ref = oracle.render(xyz, rgb, s0, W, H, oracle.rotation(0.0, p, 0.0), mode=0)[0]
        assert np.array_equal(frames[k], ref), f"frame {k}"


def test_shard_helpers():
    import rust_sloth_b200  # noqa: F401
    from rust_sloth_b200 import turntable as tt
    for n in (0, 1, 7, 360):
        for w in (1, 2, 3, 8):
            got = sorted(k for r in range(w) for k in tt.frame_shard(n, r, w))
            assert got == list(range(n))
    for H in (1, 40, 2160, 4321):
        for w in (1, 2, 8):
            e = tt.band_edges(H, w)
            assert e[0] == 0 and e[-1] == H and all(a <= b for a, b in zip(e, e[1:]))


def test_webify_stream_framing():
    import rust_sloth_b200  # noqa: F401
    from rust_sloth_b200 import turntable as tt
    f = [np.array([ord("@") | 7 << 8], np.uint32), np.array([ord(" ")], np.uint32)]
    s = tt.webify_stream(f)
    assert s == (b'let frames = [\n`\n<span style="color:rgb(7,0,0)">@`,\n`\n<span style="color:rgb(0,0,0)"> `];\n')


def test_webify_framing_matches_the_reference_player_data():
    """the -j wire format.  src-webify/data.js (a stale 100-frame export shipped with the
    reference; its spans are run-length merged, so only the framing is comparable) and our stream use the same
    tokens: `let frames = [`, a back-tick line before every frame, "`," after it, "`];" after the last."""
    import re
    import rust_sloth_b200  # noqa: F401
    from rust_sloth_b200 import turntable as tt
    cells = np.full(10 * 4 + 4, ord(" "), np.uint32)
    cells[1::10][:4] = ord("\n")
    ours = tt.webify_stream([cells] * 3).decode()
    assert ours.startswith("let frames = [\n`\n<span") and ours.endswith("`];\n")
    assert len(re.findall(r"(?m)^`$", ours)) == 3 and ours.count("`,\n") == 2
    path = "/root/reference/src-webify/data.js"
    if os.path.exists(path):
        ref = open(path).read()
        assert ref.startswith("let frames = [\n`\n<span") and ref.endswith("`];\n")
        n = len(re.findall(r"(?m)^`$", ref))
        assert n == 100 and ref.count("`,\n") == n - 1
        # render.js only needs `frames` to be an array of strings: same JS shape on both sides
        assert re.fullmatch(r"let frames = \[\n(`\n[^`]*`,\n)*`\n[^`]*`\];\n", ours)
        assert re.fullmatch(r"let frames = \[\n(`\n[^`]*`,\n)*`\n[^`]*`\];\n", ref)


def test_host_band_renderer_edges_and_shared_frame():
    """multigpu.HostBandRenderer without a GPU: the shared-memory frame, the per-rank counters and the work-balanced
    band edges (a stub context writes its band index into its rows)."""
    import numpy as np
    from rust_sloth_b200 import multigpu

    class StubCtx:
        def __init__(self, r): self.r, self.band = r, None
        def resize(self, w, h): self.w, self.h = w, h
        def set_band(self, a, b): self.band = (a, b)
        def render_into(self, rot, out): out[:(self.band[1] - self.band[0]) * self.w] = 100 + self.r

    W, H, world = 12, 40, 4
    name = f"sloth_test_{os.getpid()}"
    rs_ = [multigpu.HostBandRenderer(StubCtx(r), W, H, r, world, name, page_lock=False) for r in range(world)]
    try:
        assert rs_[0].edges == [0, 10, 20, 30, 40]
        for r in rs_:
            r.render(None)
        frame = rs_[2].wait(1)
        body = frame[:W * H].reshape(H, W)
        assert all((body[10 * r:10 * r + 10] == 100 + r).all() for r in range(world))
        assert (frame[W * H:] == ord(" ")).all()
        # band 1 did 3x the work of the others: its rows shrink, edges stay strictly increasing and end at H
        for r in rs_:
            edges = r.rebalance([10, 30, 10, 10])
        assert edges[0] == 0 and edges[-1] == H and all(a < b for a, b in zip(edges, edges[1:]))
        assert edges[2] - edges[1] < 10 and edges == rs_[3].edges
        for r in rs_:
            r.render(None)
        body = rs_[0].wait(2)[:W * H].reshape(H, W)
        assert all((body[edges[r]:edges[r + 1]] == 100 + r).all() for r in range(world))
    finally:
        for r in reversed(rs_):
            r.barrier = lambda: None
            r.close()
