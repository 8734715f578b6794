"""C1 (examples/simulate_gaussian_source.py, 120^3, two full-volume energy videos every 3rd step):
per-step cost of the field update alone vs with detector accumulation, forward and time-reversed,
for the row-marching detector kernels (FDTDX_B200_DET_VOLUME=1, default) and the generic ones (=0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import configs
import fdtdx_b200 as fx
from fdtdx_b200.fdtd import get_plan


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us per step


for pml in (True, False):
    for vol in ("1", "0"):
        os.environ["FDTDX_B200_DET_VOLUME"] = vol
        objects, arrays, cfg = configs.build_c1(pml=pml)
        dev = arrays.to_torch("cuda")
        plan = get_plan(dev, objects, cfg)
        T = cfg.time_steps_total
        plan.run_forward(0, 6, True, True, True)
        n = 300
        us_field = timed(lambda: plan.run_forward(6, n, False, True, True), n)
        l0 = plan.launch_count()
        us_det = timed(lambda: plan.run_forward(6 + n, 198, True, True, True), 198)
        launches = (plan.launch_count() - l0) / 198
        active = 198 / 3
        per_active = (us_det - us_field) * 198 / active
        us_bwd = timed(lambda: plan.run_reverse(T, 150, False, True), 150) if pml else float("nan")
        us_bwd_det = timed(lambda: plan.run_reverse(T - 150, 150, True, True), 150) if pml else float("nan")
        print(f"C1 {'CPML' if pml else 'periodic'} det_volume={vol}: field update {us_field:.1f} us/step, with 2 energy videos {us_det:.1f} us/step "
              f"({launches:.2f} launches/step) -> detector cost {per_active:.1f} us per active step = {per_active / us_field:.2f} x field update; "
              f"reverse {us_bwd:.1f} us/step, reverse + inverse video {us_bwd_det:.1f} us/step", flush=True)
        objects.__dict__.pop("_plan_cache", None)
