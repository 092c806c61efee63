"""Test scenes: the reference's bundled models as committed soups (tests/golden/*.npz,
written by tests/golden/make_golden.py) plus synthetic ones.  Nothing here reads /root/reference."""
import functools
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
PI = float(np.float32(np.pi))


@functools.lru_cache(maxsize=None)
def _npz(name):
    return np.load(os.path.join(GOLDEN, name))


@functools.lru_cache(maxsize=None)
def soup(scene: str):
    """-> (xyz (n,9) f32, rgb (n,3) u8, scale0 f32)"""
    if scene == "suzy_suzy":  # "models/suzy.obj models/suzy.obj": the same mesh queue twice
        xyz, rgb, s0 = soup("suzy")
        return np.concatenate([xyz, xyz]), np.concatenate([rgb, rgb]), s0
    z = _npz("hand.npz" if scene == "hand" else "models.npz")
    return z[scene + "_xyz"], z[scene + "_rgb"], np.float32(z[scene + "_scale0"])


def mesh_sizes(scene: str):
    z = _npz("hand.npz" if scene == "hand" else "models.npz")
    return z[scene + "_sizes"]


@functools.lru_cache(maxsize=None)
def golden():
    with open(os.path.join(GOLDEN, "oracle_frames.json")) as f:
        return json.load(f)
