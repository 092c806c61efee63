"""Host-side frame loop of `sloth ... image -j N` (src/main.rs:55-58,85-106) and the
multi-GPU work split.  Frames of a turntable depend only on their angle
(main.rs:92-96), so they shard over GPUs with no collective; one huge frame shards by
destination-row bands that are gathered once."""
from __future__ import annotations

import numpy as np

from . import flush_bytes, rotation_from_euler, turntable_pitches


def frame_shard(n_frames: int, rank: int, world: int) -> list[int]:
    """Frames rendered by `rank`: k = rank, rank+world, ... (round-robin keeps ranks in step)."""
    return list(range(rank, n_frames, world))


def band_edges(height: int, world: int) -> list[int]:
    """Row bands [e[i], e[i+1]) of one frame, as equal as integer rows allow."""
    return [height * i // world for i in range(world + 1)]


def turntable_rotations(x: float, y: float, z: float, n_frames: int) -> np.ndarray:
    """(F,16) column-major rotations of every frame the reference would render."""
    pitches = turntable_pitches(y, n_frames)
    return np.stack([rotation_from_euler(x, p, z) for p in pitches])


def assemble_bands(parts, width: int, height: int, image: bool) -> np.ndarray:
    """Concatenate band outputs (in row order) and append the image-mode tail (context.rs:38-39)."""
    cells = np.concatenate([np.asarray(p, np.uint32).reshape(-1) for p in parts])
    assert cells.size == width * height
    if image:
        cells = np.concatenate([cells, np.full(height, ord(" "), np.uint32)])
    return cells


def webify_stream(frames, color: bool = True) -> bytes:
    """The complete stdout of `image -j N` for already rendered frames (main.rs:55-57,85-87,98-104)."""
    out = bytearray(b"let frames = [\n")
    n = len(frames)
    for k, cells in enumerate(frames):
        out += b"`\n"
        out += flush_bytes(cells, color, True, True)
        out += b"`];\n" if k == n - 1 else b"`,\n"
    return bytes(out)


def interleave_shards(shards, n_frames: int, world: int):
    """Inverse of frame_shard: shards[r][i] is frame r + i*world."""
    frames = [None] * n_frames
    for r in range(world):
        for i, k in enumerate(frame_shard(n_frames, r, world)):
            frames[k] = shards[r][i]
    return frames
