"""Standard perf lines for kernel work (not the bench): prints Gcell/s for a fixed set of shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_perf import run

if __name__ == "__main__":
    run((512, 512, 512), boundaries="periodic")
    run((512, 512, 512))
    run((1024, 1024, 512))
    run((1897, 291, 128), thickness=12, nonuniform=True)
    run((1897, 291, 128), thickness=12)
