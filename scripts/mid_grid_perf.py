import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from quick_perf import run
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(120, 120, 120), (135, 135, 76), (192, 192, 128), (256, 256, 256), (384, 384, 128), (948, 145, 68)]
for shape in shapes:
    run(shape, steps=200, thickness=10)
