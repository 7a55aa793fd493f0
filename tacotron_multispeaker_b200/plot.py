"""Alignment image writer with the call shape of the reference's
``util/plot.py:6-20`` (``plot_alignment(alignment, path, info=None)``).

matplotlib is not a dependency here: the alignment ``[encoder steps, decoder
steps]`` is written as a colour-mapped PNG (origin lower-left, nearest
neighbour, like ``imshow(origin='lower', interpolation='none')``); axes,
colour bar and the ``info`` caption of the reference figure are not drawn --
``info`` goes into a PNG text chunk.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

# anchor colours of a viridis-like map (dark violet -> blue -> green -> yellow)
_ANCHORS = np.array([[68, 1, 84], [59, 82, 139], [33, 145, 140], [94, 201, 98], [253, 231, 37]], np.float64)


def _colormap(x: np.ndarray) -> np.ndarray:
    x = np.clip(x, 0.0, 1.0) * (len(_ANCHORS) - 1)
    i = np.minimum(x.astype(np.int64), len(_ANCHORS) - 2)
    f = (x - i)[..., None]
    return (_ANCHORS[i] * (1 - f) + _ANCHORS[i + 1] * f + 0.5).astype(np.uint8)


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def plot_alignment(alignment, path, info=None, min_size=(640, 480)):
    a = np.asarray(alignment, np.float64)
    if a.ndim != 2 or a.size == 0:
        raise ValueError("alignment must be a non-empty [encoder steps, decoder steps] matrix")
    lo, hi = float(a.min()), float(a.max())
    a = (a - lo) / (hi - lo) if hi > lo else np.zeros_like(a)
    a = a[::-1]                                            # origin='lower'
    sy = max(1, -(-min_size[1] // a.shape[0]))
    sx = max(1, -(-min_size[0] // a.shape[1]))
    rgb = _colormap(np.repeat(np.repeat(a, sy, axis=0), sx, axis=1))
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
    text = "x: Decoder timestep, y: Encoder timestep" + ("; " + info if info else "")
    png += _chunk(b"tEXt", b"Comment\x00" + text.encode("latin-1", "replace"))
    png += _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)
