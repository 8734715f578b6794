import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import numpy as np, torch
import fdtdx_b200 as fx
from fdtdx_b200.fdtd import get_plan
from oracle import yee_torch
from scenes import build_scene, rel_l2
for kw in (dict(shape=(16,14,20), thickness=4), dict(shape=(16,14,20), thickness=4, kappa=True), dict(shape=(16,14,20), thickness=4, boundaries={"min_x":"pml","max_x":"pml","min_y":"periodic","max_y":"periodic","min_z":"periodic","max_z":"periodic"})):
    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build_scene(recorder=rec, source="plane_z", time=3e-15, **kw)
    T = cfg.time_steps_total
    E0 = torch.zeros(arrays.fields.E.shape, dtype=torch.float64, requires_grad=True)
    H0 = torch.zeros(arrays.fields.H.shape, dtype=torch.float64, requires_grad=True)
    ie = torch.tensor(arrays.inv_permittivities.astype(np.float64), requires_grad=True)
    E, H, det = yee_torch.run_forward(arrays.reset(), objects, cfg, T, inv_eps=ie, dtype=torch.float64, E0=E0, H0=H0)
    loss = 0.5*(E*E).sum() + 0.5*(H*H).sum()
    loss.backward()
    dev = arrays.to_torch("cuda")
    t_end, out = fx.run_fdtd(dev, objects, cfg)
    work = out.aset("fields->E", out.fields.E.clone()).aset("fields->H", out.fields.H.clone())
    cotE, cotH = out.fields.E.clone(), out.fields.H.clone()
    g = torch.zeros_like(work.inv_permittivities)
    plan = get_plan(work, objects, cfg)
    plan.run_adjoint(work, T, T, cotE, cotH, {}, g, None)
    torch.cuda.synchronize()
    m = 5
    inner = (slice(None), slice(m,-m), slice(m,-m), slice(m,-m))
    print(kw.get("kappa"), kw.get("boundaries") is not None, "lamE0 full", rel_l2(cotE.cpu().numpy(), E0.grad.numpy()), "inner", rel_l2(cotE.cpu().numpy()[inner], E0.grad.numpy()[inner]),
          "lamH0 full", rel_l2(cotH.cpu().numpy(), H0.grad.numpy()), "g full", rel_l2(g.cpu().numpy(), ie.grad.numpy()), "g inner", rel_l2(g.cpu().numpy()[inner], ie.grad.numpy()[inner]),
          "recon E0 max", float(work.fields.E.abs().max()), "E_T max", float(out.fields.E.abs().max()))
