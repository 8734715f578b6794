"""Time the reversible_fdtd forward (with PML-interface recording) and backward (time-reversed
reconstruction + fused adjoint) passes on a C4-style inverse-design scene.  Not the bench."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import argparse
import numpy as np, torch
import fdtdx_b200 as fx
from scenes import build_scene

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="192,192,128")
ap.add_argument("--thickness", type=int, default=10)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--bf16", action="store_true", help="record the interfaces in bf16 like the reference's default recorder modules")
a = ap.parse_args()
shape = tuple(int(v) for v in a.shape.split(","))
rec = fx.Recorder(modules=[])
objects, arrays, cfg = build_scene(shape=shape, thickness=a.thickness, source="plane_z", detectors=("poynting", "phasor"), recorder=rec, time=1e-12)
dt = cfg.time_step_duration
cfg = cfg.aset("time", a.steps * dt * 1.0001)
objects, arrays, cfg = build_scene(shape=shape, thickness=a.thickness, source="plane_z", detectors=("poynting", "phasor"), recorder=rec, time=a.steps * dt * 1.0001)
T = cfg.time_steps_total
cells = float(np.prod(shape))
dev = arrays.to_torch("cuda")
dev.inv_permittivities.requires_grad_(True)
for rep in range(2):
    dev.inv_permittivities.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, out = fx.run_fdtd(dev, objects, cfg)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss = sum((v.real**2 + v.imag**2).sum() if v.is_complex() else v.sum() * 1e18 for st in out.detector_states.values() for v in st.values())
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"shape={shape} T={T}: forward+record {1e3*(t1-t0)/T:.3f} ms/step ({cells*T/(t1-t0)/1e9:.1f} Gcell/s), "
          f"backward (reverse + adjoint) {1e3*(t2-t1)/T:.3f} ms/step ({cells*T/(t2-t1)/1e9:.1f} Gcell/s), |grad| {float(dev.inv_permittivities.grad.abs().max()):.3e}", flush=True)
