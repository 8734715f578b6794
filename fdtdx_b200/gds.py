"""Minimal GDSII stream reader + polygon rasteriser for the C2 directional-coupler scene.

The reference imports ``performance/coupler.gds`` with gdstk and rasterises it through its
``gds_layer_stack`` objects (``performance/directional_coupler.py:117-140, 210-225``); neither gdstk nor
that object machinery is on the hot path.  What the Yee step needs from the file is a boolean core
mask, so this module reads the BOUNDARY polygons of one cell / layer straight from the record stream
(GDSII: 2-byte length, record type, data type, big-endian payload; 8-byte excess-64 reals in UNITS)
and marks the grid cells whose centre lies inside a polygon (even-odd rule).
"""

from __future__ import annotations

import struct

import numpy as np

_BGNSTR, _STRNAME, _ENDSTR, _BOUNDARY, _LAYER, _XY, _ENDEL, _UNITS = 0x05, 0x06, 0x07, 0x08, 0x0D, 0x10, 0x11, 0x03


def _real8(b: bytes) -> float:
    sign = -1.0 if b[0] & 0x80 else 1.0
    exp = (b[0] & 0x7F) - 64
    mant = int.from_bytes(b[1:8], "big") / float(1 << 56)
    return sign * mant * 16.0**exp


def read_gds_polygons(path: str, cell: str, layer: int) -> list[np.ndarray]:
    """Polygons (N, 2) float64 in METRES of every BOUNDARY on ``layer`` in structure ``cell``."""
    data = open(path, "rb").read()
    pos, unit_m, cur, in_boundary, lay, polys = 0, 1e-9, None, False, None, []
    while pos + 4 <= len(data):
        ln, rt, _dt = struct.unpack(">HBB", data[pos:pos + 4])
        if ln < 4:
            break
        payload = data[pos + 4:pos + ln]
        pos += ln
        if rt == _UNITS:
            unit_m = _real8(payload[8:16])  # metres per database unit
        elif rt == _STRNAME:
            cur = payload.rstrip(b"\0").decode()
        elif rt == _ENDSTR:
            cur = None
        elif rt == _BOUNDARY:
            in_boundary, lay = True, None
        elif rt == _LAYER and in_boundary:
            lay = struct.unpack(">h", payload)[0]
        elif rt == _XY and in_boundary and cur == cell and lay == layer:
            xy = np.array(struct.unpack(f">{len(payload) // 4}i", payload), np.float64).reshape(-1, 2) * unit_m
            polys.append(xy[:-1] if np.array_equal(xy[0], xy[-1]) else xy)
        elif rt == _ENDEL:
            in_boundary = False
    return polys


def rasterize_polygons(polys, xc: np.ndarray, yc: np.ndarray) -> np.ndarray:
    """(len(xc), len(yc)) bool: cell centre inside any polygon (even-odd crossing count along +x)."""
    X, Y = np.meshgrid(np.asarray(xc, np.float64), np.asarray(yc, np.float64), indexing="ij")
    inside = np.zeros(X.shape, bool)
    for p in polys:
        x0, y0 = p[:, 0], p[:, 1]
        x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
        lo, hi = p.min(axis=0), p.max(axis=0)
        ix = np.nonzero((xc >= lo[0]) & (xc <= hi[0]))[0]
        iy = np.nonzero((yc >= lo[1]) & (yc <= hi[1]))[0]
        if ix.size == 0 or iy.size == 0:
            continue
        sx, sy = slice(ix[0], ix[-1] + 1), slice(iy[0], iy[-1] + 1)
        px, py = X[sx, sy][..., None], Y[sx, sy][..., None]
        cond = (y0 > py) != (y1 > py)
        with np.errstate(divide="ignore", invalid="ignore"):
            xint = x0 + (py - y0) * (x1 - x0) / (y1 - y0)
        cross = cond & (px < xint)
        inside[sx, sy] ^= (cross.sum(axis=-1) % 2).astype(bool)
    return inside
