"""Test-side formatter of Context::flush's `-j` cell stream (src/context.rs:78-87), independent of the product's
Python layer: `<span style="color:rgb(r,g,b)">c` per cell, never closed.  Vectorised over the distinct cell values
of a frame (a 1080p frame has 2 M cells but a few hundred distinct ones)."""
import numpy as np


def webify_cells(cells: np.ndarray) -> bytes:
    cells = np.asarray(cells, np.uint32)
    uniq, inv = np.unique(cells, return_inverse=True)
    table = [b'<span style="color:rgb(%d,%d,%d)">' % ((int(c) >> 8) & 255, (int(c) >> 16) & 255, (int(c) >> 24) & 255) + bytes([int(c) & 255])
             for c in uniq]
    return b"".join(table[i] for i in inv.reshape(-1))


def webify_length(cells: np.ndarray) -> int:
    """Length of webify_cells(cells) without building it."""
    cells = np.asarray(cells, np.uint32)
    digits = np.ones(256, np.int64)
    digits[10:] = 2
    digits[100:] = 3
    r, g, b = (cells >> 8) & 255, (cells >> 16) & 255, (cells >> 24) & 255
    return int((len(b'<span style="color:rgb(,,)">') + 1) * cells.size + digits[r].sum() + digits[g].sum() + digits[b].sum())
