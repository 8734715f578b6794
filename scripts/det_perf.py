"""Per-step cost of detector accumulation on a C1-style scene (120^3, full-volume energy slices + phasor plane)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from scenes import build_scene
from fdtdx_b200.fdtd import get_plan

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "120,120,120").split(","))
for dets in ((), ("energy_slices",), ("energy_slices", "phasor", "poynting"), ("energy",)):
    objects, arrays, cfg = build_scene(shape=shape, thickness=10, source="plane_z", detectors=dets, time=2e-13)
    dev = arrays.to_torch("cuda")
    plan = get_plan(dev, objects, cfg)
    plan.run_forward(0, 5, True, False, True)
    torch.cuda.synchronize()
    n = min(200, cfg.time_steps_total - 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run_forward(5, n, True, False, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"shape={shape} detectors={dets}: {ms*1e3:.1f} us/step ({np.prod(shape)/ms/1e6:.1f} Gcell/s)", flush=True)
    objects.__dict__.pop("_plan_cache", None)
