import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
from quick_perf import run
for shape in ((120, 120, 120), (135, 135, 76), (192, 192, 128), (256, 256, 256), (384, 384, 128), (948, 145, 68)):
    run(shape, steps=200, thickness=10)
