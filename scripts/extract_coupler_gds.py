"""Extracts the layer-1 BOUNDARY polygons of cell "coupler" from the reference's
performance/coupler.gds into fdtdx_b200/data/coupler_gds.npz (run in the build container, where
/root/reference exists; the GPU box only sees the derived fixture)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fdtdx_b200.gds import read_gds_polygons

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/performance/coupler.gds"
polys = read_gds_polygons(src, "coupler", 1)
out = os.path.join(ROOT, "fdtdx_b200", "data", "coupler_gds.npz")
np.savez_compressed(out, n=len(polys), **{f"p{i}": p for i, p in enumerate(polys)})
for p in polys:
    print(p.shape, p.min(axis=0) * 1e6, p.max(axis=0) * 1e6)
print("wrote", out, os.path.getsize(out), "bytes")
