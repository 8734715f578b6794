"""A few C1 steps (120^3 CPML, two energy videos) for ncu: 12 forward steps, then 6 reverse."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import configs
from fdtdx_b200.fdtd import get_plan

objects, arrays, cfg = configs.build_c1(pml=(os.environ.get("C1_PML", "1") == "1"))
dev = arrays.to_torch("cuda")
plan = get_plan(dev, objects, cfg)
plan.run_forward(0, 12, True, True, True)
torch.cuda.synchronize()
if os.environ.get("C1_REVERSE", "1") == "1":
    plan.run_reverse(12, 6, True, True)
torch.cuda.synchronize()
print("done")
