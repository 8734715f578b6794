"""Time BASELINE config C4 (examples/optimize_ceviche_corner.py: (135,135,75), 1311 steps, recorder
[every-5, fp8_e4m3fnuz], gated flux detectors, loss = -flux_out / flux_in) forward and backward through
``run_fdtd`` + ``loss.backward()``.  Not the bench."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import fdtdx_b200 as fx
import configs
from make_config_golden import c4_loss

objects, arrays, cfg = configs.build_c4()
T = cfg.time_steps_total
shape = objects.volume.grid_shape
cells = float(np.prod(shape))
dev = arrays.to_torch("cuda")
dev.inv_permittivities.requires_grad_(True)
for rep in range(int(os.environ.get("REPS", "3"))):
    dev.inv_permittivities.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, out = fx.run_fdtd(dev, objects, cfg)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss = c4_loss(out.detector_states)
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"C4 {shape} T={T}: forward+record {1e3*(t1-t0)/T:.3f} ms/step ({cells*T/(t1-t0)/1e9:.1f} Gcell/s), "
          f"backward (reverse + adjoint) {1e3*(t2-t1)/T:.3f} ms/step ({cells*T/(t2-t1)/1e9:.1f} Gcell/s) = {(t2-t1)/(t1-t0):.2f} x forward, "
          f"loss {float(loss):.6e}, |grad| {float(dev.inv_permittivities.grad.abs().max()):.3e}", flush=True)
