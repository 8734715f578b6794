"""Throughput of the material tiers on seeded scenes (not the bench): iso / diag / sigma / ADE / full tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from scenes import build_scene, seed_fields
from fdtdx_b200.fdtd import get_plan

def run(name, steps=30, **kw):
    objects, arrays, cfg = build_scene(time=1e-12, **kw)
    seed_fields(arrays, seed=1)
    dev = arrays.to_torch("cuda")
    plan = get_plan(dev, objects, cfg)
    plan.run_forward(0, 3, False, False, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run_forward(3, steps, False, False, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cells = float(np.prod(kw["shape"]))
    print(f"{name:28s} shape={kw['shape']} {ms:.3f} ms/step  {cells/ms/1e6:.2f} Gcell/s", flush=True)
    objects.__dict__.pop("_plan_cache", None)

if __name__ == "__main__":
    sh = (256, 256, 256)
    run("iso", shape=sh, thickness=10)
    run("diag eps", shape=sh, thickness=10, eps_tier=3)
    run("diag eps+sigma+mu", shape=sh, thickness=10, eps_tier=3, sigma_E=True, mu_tier=3, sigma_H=True)
    run("ADE 1 pole", shape=sh, thickness=10, poles=1)
    run("ADE 2 poles c4 diag", shape=(192, 192, 192), thickness=10, poles=2, c4=True, eps_tier=3, coeff_tier=3)
    run("full tensor eps9", shape=(192, 192, 192), thickness=10, eps_tier=9)
    run("full tensor eps9 mu9", shape=(192, 192, 192), thickness=10, eps_tier=9, mu_tier=9)
    run("C3b 120^3 eps9", shape=(120, 120, 120), thickness=10, eps_tier=9, steps=100)
