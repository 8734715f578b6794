"""Quick kernel-throughput probe (not the bench): vacuum/dielectric box + CPML, isotropic."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fdtdx_b200 as fx
from fdtdx_b200.fdtd import get_plan

def run(shape, steps=20, boundaries="pml", thickness=10, xchunk=0, rows=0, nonuniform=False):
    if nonuniform:
        edges = [np.concatenate([[0.0], np.cumsum(50e-9 * (1 + 0.5 * np.random.default_rng(a).random(n)))]) for a, n in enumerate(shape)]
        grid = fx.RectilinearGrid(*edges)
    else:
        grid = fx.UniformGrid(spacing=50e-9)
    cfg = fx.SimulationConfig(time=1e-12, grid=grid)
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = fx.boundary_objects_from_config(shape, cfg, boundaries, thickness=thickness)
    objects = fx.ObjectContainer([vol, *bl])
    dev = torch.device("cuda")
    f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
    psiE = {p.name: (f(*p.grid_shape), f(*p.grid_shape)) for p in objects.pml_objects}
    psiH = {p.name: (f(*p.grid_shape), f(*p.grid_shape)) for p in objects.pml_objects}
    E = 1e-3 * torch.randn(3, *shape, device=dev); H = 1e-3 * torch.randn(3, *shape, device=dev)
    eps = torch.ones(1, *shape, device=dev)
    arrays = fx.ArrayContainer(fields=fx.FieldState(E=E, H=H, psi_E=psiE, psi_H=psiH), inv_permittivities=eps, inv_permeabilities=1.0, detector_states={}, recording_state=None)
    plan = get_plan(arrays, objects, cfg)
    plan.set_tuning(xchunk, rows)
    plan.run_forward(0, 3, False, False, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run_forward(3, steps, False, False, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cells = float(np.prod(shape))
    gc = cells / ms / 1e6
    print(f"shape={shape} bnd={boundaries} nonuni={nonuniform} xchunk={xchunk} rows={rows}: {ms:.3f} ms/step  {gc:.2f} Gcell/s  {gc*76/6513.2*100:.1f}% of HBM roofline (76 B/cell)", flush=True)
    del plan
    objects.__dict__.pop("_plan_cache", None)

if __name__ == "__main__":
    run((512, 512, 512))
    run((512, 512, 512), boundaries="periodic")
    for xc in (8, 16, 32, 64, 128):
        run((512, 512, 512), xchunk=xc)
    for rows in (2, 4):
        run((512, 512, 512), rows=rows)
    run((1024, 1024, 512))
    run((1897, 291, 128), thickness=12, nonuniform=True)
    run((1897, 291, 128), thickness=12)
