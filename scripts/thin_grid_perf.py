"""Thin / ragged grids: staged half-steps with flat tiles (default) vs the 64- / 128-cell tile rows (FDTDX_B200_TMA_FLAT=0)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r + "/scripts")
from quick_perf import run
for shape in ((948, 145, 65), (948, 145, 68), (948, 145, 64), (512, 512, 96), (512, 512, 40), (135, 135, 75), (1897, 291, 128)):
    run(shape, steps=100, thickness=10)
run((948, 145, 65), steps=100, thickness=10, nonuniform=True)
''' % (ROOT, ROOT)
for v in ({}, {"FDTDX_B200_TMA_FLAT": "0"}):
    print(v or "default (flat tiles where they apply)", flush=True)
    env = dict(os.environ); env.update(v)
    subprocess.run([sys.executable, "-c", CODE], env=env)
