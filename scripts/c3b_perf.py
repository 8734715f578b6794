import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, configs
from fdtdx_b200.fdtd import get_plan
objects, arrays, cfg = configs.build_c3b()
dev = arrays.to_torch("cuda")
for xc, rows in ((0, 0), (2, 0), (4, 0), (8, 0), (4, 8), (8, 8), (16, 8)):
    objects.__dict__.pop("_plan_cache", None)
    plan = get_plan(dev, objects, cfg)
    plan.set_tuning(xc, rows)
    plan.run_forward(0, 6, False, False, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record(); plan.run_forward(6, n, False, False, True); e1.record(); torch.cuda.synchronize()
    print(f"C3b (periodic x,y; nine-component slab) xchunk={xc} rows={rows}: {e0.elapsed_time(e1)/n*1e3:.1f} us/step", flush=True)
