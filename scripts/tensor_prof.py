"""A few steps of the full-tensor tier (192^3, eps 9-component, 10-cell CPML) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from scenes import build_scene, seed_fields
from fdtdx_b200.fdtd import get_plan

objects, arrays, cfg = build_scene(time=1e-12, shape=(192, 192, 192), thickness=10, eps_tier=9, mu_tier=int(os.environ.get("MU_TIER", "0")))
seed_fields(arrays, seed=1)
dev = arrays.to_torch("cuda")
plan = get_plan(dev, objects, cfg)
plan.run_forward(0, int(os.environ.get("STEPS", "6")), False, False, True)
torch.cuda.synchronize()
print("done")
