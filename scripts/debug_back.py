import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import fdtdx_b200 as fx
from oracle import yee
from scenes import build_scene, rel_l2
rec = fx.Recorder(modules=[])
objects, arrays, cfg = build_scene(source="plane_z", detectors=("energy_slices", "inverse_energy"), recorder=rec, time=6e-15)
T = cfg.time_steps_total
st_o = yee.checkpointed_fdtd(arrays, objects, cfg)
dev = arrays.to_torch("cuda")
st_g = fx.run_fdtd(dev, objects, cfg)
print("fwd", rel_l2(st_g[1].fields.E.cpu().numpy(), st_o[1].fields.E))
for k, ref in st_o[1].recording_state.data.items():
    print(k, rel_l2(st_g[1].recording_state.data[k].cpu().numpy(), ref), ref.shape)
for i in range(6):
    st_o = yee.backward(st_o, cfg, objects, None, True, True)
    st_g = fx.backward(st_g, cfg, objects, None, True, True)
    Eg = st_g[1].fields.E.cpu().numpy(); Hg = st_g[1].fields.H.cpu().numpy()
    dE = np.abs(Eg - st_o[1].fields.E); dH = np.abs(Hg - st_o[1].fields.H)
    print(st_o[0], st_g[0], "relE", rel_l2(Eg, st_o[1].fields.E), "relH", rel_l2(Hg, st_o[1].fields.H), "argmax dE", np.unravel_index(dE.argmax(), dE.shape), dE.max(), "argmax dH", np.unravel_index(dH.argmax(), dH.shape), dH.max(), np.abs(st_o[1].fields.E).max())
