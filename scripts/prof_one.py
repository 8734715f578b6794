"""Run a few steps of one configuration (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import scripts.quick_perf as qp
ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="512,512,512"); ap.add_argument("--bnd", default="pml"); ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--xchunk", type=int, default=0); ap.add_argument("--rows", type=int, default=0); ap.add_argument("--nonuniform", action="store_true"); ap.add_argument("--thickness", type=int, default=10)
a = ap.parse_args()
qp.run(tuple(int(v) for v in a.shape.split(",")), steps=a.steps, boundaries=a.bnd, xchunk=a.xchunk, rows=a.rows, nonuniform=a.nonuniform, thickness=a.thickness)
