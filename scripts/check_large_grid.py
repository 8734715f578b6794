"""64-bit indexing check: a grid with 3*N > 2^31 elements per field (1024 x 1024 x 768 = 805 Mcell, 9.7 GB
per field array).  Staged and marching kernels must agree bit-for-bit after a few steps from a seeded
state, and energy must be finite and non-zero far from the origin (the last planes are touched)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fdtdx_b200 as fx
from fdtdx_b200 import workloads as W
from fdtdx_b200.fdtd import get_plan

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1024,1024,768").split(","))
dev = torch.device("cuda")
outs = []
for tma in (1, 0):
    objects, arrays, cfg = W.build_box(shape, device=dev, thickness=10)
    g = torch.Generator(device=dev); g.manual_seed(1)
    arrays.fields.E.copy_(1e-3 * torch.randn(arrays.fields.E.shape, device=dev, generator=g))
    arrays.fields.H.copy_(1e-3 * torch.randn(arrays.fields.H.shape, device=dev, generator=g))
    plan = get_plan(arrays, objects, cfg)
    plan.set_tma(tma, 0)
    plan.run_forward(0, 3, False, False, True)
    torch.cuda.synchronize()
    E, H = arrays.fields.E, arrays.fields.H
    chk = [float(E[c, -40:].double().pow(2).sum()) for c in range(3)] + [float(H[c, -40:].double().pow(2).sum()) for c in range(3)]
    tot = float(plan.total_energy(arrays).item())
    outs.append((E.clone() if tma else None, H.clone() if tma else None, chk, tot, E, H))
    print(f"tma={tma} shape={shape} cells={np.prod(shape)/1e6:.0f}M tail-plane sums {['%.6e' % v for v in chk]} total energy {tot:.6e}", flush=True)
    if tma == 0:
        same = torch.equal(outs[0][0], E) and torch.equal(outs[0][1], H)
        print("LARGE GRID CHECK", "PASSED" if same and all(np.isfinite(chk)) and min(chk) > 0 else "FAILED", flush=True)
    objects.__dict__.pop("_plan_cache", None)
