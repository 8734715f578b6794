import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
from quick_perf import run
for shape in ((120, 120, 120), (192, 192, 128), (256, 256, 256), (384, 384, 384), (512, 512, 512)):
    for xc in (0, 16):
        run(shape, steps=100, boundaries="periodic", xchunk=xc)
