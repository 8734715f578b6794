"""Generates the config-faithful golden files ``tests/golden/cfg_*.npz`` from the CPU oracle.

    python tests/golden/make_config_golden.py c1 | c3a | c3b | c4 | c2        (minutes each, CPU only)

Each file holds what ``tests/test_config_parity.py`` compares against the CUDA path after running the
SAME BASELINE.json configuration at its real size and step count (tests/configs.py): detector
states (sub-sampled where they are volume-sized), sub-sampled final fields plus whole-array norms,
and - for C4 - the time-reversal gradient on the design block.  These are ORACLE outputs (the reference
itself cannot run in this image): "parity unpinned against reference outputs".
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import configs  # noqa: E402
from configs import field_norms, sub, t_pick  # noqa: E402
from oracle import yee  # noqa: E402


def _fields(prefix, arrays, out):
    out[f"{prefix}_E"] = sub(arrays.fields.E)
    out[f"{prefix}_H"] = sub(arrays.fields.H)
    out[f"{prefix}_E_norm"] = field_norms(arrays.fields.E)
    out[f"{prefix}_H_norm"] = field_norms(arrays.fields.H)


def _log(name, t0, t, T):
    if t % 50 == 0:
        print(f"[{name}] step {t}/{T}  {time.time() - t0:.0f}s", flush=True)


def run_forward(name, objects, arrays, cfg, record_boundaries):
    st = (0, arrays.reset())
    T, t0 = cfg.time_steps_total, time.time()
    while st[0] < T:
        st = yee.forward(st, cfg, objects, None, True, record_boundaries, True)
        _log(name, t0, st[0], T)
    return st


def run_backward(name, st, objects, cfg, until=0):
    t0 = time.time()
    while st[0] > until:
        st = yee.backward(st, cfg, objects, None, True, True)
        _log(name + "-bwd", t0, st[0], cfg.time_steps_total)
    return st


def energy_video(prefix, state, out, n_frames):
    for key in ("XY Plane", "XZ Plane", "YZ Plane"):
        v = np.asarray(state[key])
        out[f"{prefix}_{key[:2]}"] = np.ascontiguousarray(v[t_pick(n_frames)][:, ::2, ::2])
        out[f"{prefix}_{key[:2]}_norm"] = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2)))  # every frame


def make_c1():
    objects, arrays, cfg = configs.build_c1()
    out = {}
    st = run_forward("c1", objects, arrays, cfg, True)
    _fields("fwd", st[1], out)
    nf = st[1].detector_states["Energy Video"]["XY Plane"].shape[0]
    energy_video("video", st[1].detector_states["Energy Video"], out, nf)
    half = cfg.time_steps_total // 2
    st = run_backward("c1", st, objects, cfg, until=half)
    _fields("mid", st[1], out)
    st = run_backward("c1", st, objects, cfg, until=0)
    _fields("bwd", st[1], out)
    energy_video("bvideo", st[1].detector_states["Backwards Energy Video"], out, nf)
    return out


C3A_MID = 4500  # the pulse is in the middle of the domain (between the two detectors)


def make_c3a():
    objects, arrays, cfg = configs.build_c3a()
    out = {}
    st = (0, arrays.reset())
    while st[0] < C3A_MID:
        st = yee.forward(st, cfg, objects, None, True, False, True)
    out["mid_E"], out["mid_H"], out["mid_P"] = st[1].fields.E.copy(), st[1].fields.H.copy(), st[1].fields.dispersive_P_curr.copy()
    while st[0] < cfg.time_steps_total:
        st = yee.forward(st, cfg, objects, None, True, False, True)
    # after 350 fs the pulse has left through the CPML: the final field is a ~1e-6 residue of the peak
    out["fwd_E"], out["fwd_H"] = st[1].fields.E, st[1].fields.H
    for n in ("pulse_trace_A", "pulse_trace_B"):
        out[n] = st[1].detector_states[n]["fields"]
    return out


def make_c3b():
    objects, arrays, cfg = configs.build_c3b()
    out = {}
    st = run_forward("c3b", objects, arrays, cfg, True)
    _fields("fwd", st[1], out)
    v = st[1].detector_states["Electric Field Video"]["fields"]
    out["video"] = sub(v[t_pick(v.shape[0])])
    out["video_norm"] = np.sqrt((v.astype(np.float64).reshape(v.shape[0], -1) ** 2).sum(axis=1))
    st = run_backward("c3b", st, objects, cfg, until=cfg.time_steps_total - 150)
    _fields("mid", st[1], out)
    return out


def c4_loss(det):
    """loss = -flux_out / flux_in  (optimize_ceviche_corner.py:344-417, summed over the gated period)."""
    return -det["out flux"]["poynting_flux"].sum() / det["in flux"]["poynting_flux"].sum()


def make_c4():
    import torch

    from oracle import yee_torch

    objects, arrays, cfg = configs.build_c4()
    out = {}
    st = run_forward("c4", objects, arrays, cfg, True)
    _fields("fwd", st[1], out)
    for n in ("in flux", "out flux"):
        out[n.replace(" ", "_")] = st[1].detector_states[n]["poynting_flux"]
    for k, v in st[1].detector_states["energy_last_step"].items():
        out["energy_" + k[:2]] = v
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    t0 = time.time()
    g_eps, _ = yee_torch.reversible_gradient(st[1], objects, cfg, lambda E, H, det: c4_loss(det), progress=lambda t: _log("c4-grad", t0, t, cfg.time_steps_total))
    shape, dev, _, _ = configs.c4_geometry()
    out["grad_design"] = g_eps.numpy()[:, dev[0][0]:dev[0][1], dev[1][0]:dev[1][1], dev[2][0]:dev[2][1]].astype(np.float32)
    out["grad_norm"] = np.sqrt((g_eps.numpy().astype(np.float64)[:, 10:-10, 10:-10, 10:-10] ** 2).sum())
    out["loss"] = float(c4_loss({k: {k2: torch.as_tensor(v2) for k2, v2 in v.items()} for k, v in st[1].detector_states.items()}))
    return out


C2_MID = 2500  # the pulse is inside the coupling section


def c2_transmissions(objects, detector_states):
    """|a|^2 / |a_source|^2 of the thru and cross ports (performance/directional_coupler.py:433-444)."""
    dets = {d.name: d for d in objects.detectors}
    norm = abs(complex(dets["det_source"].compute_overlap(detector_states["det_source"])[0])) ** 2
    return {n: abs(complex(dets[n].compute_overlap(detector_states[n])[0])) ** 2 / norm for n in ("det_thru", "det_cross")}


def make_c2():
    objects, arrays, cfg = configs.build_c2(6)
    out = {}
    st = (0, arrays.reset())
    t0 = time.time()
    T = cfg.time_steps_total
    while st[0] < C2_MID:
        st = yee.forward(st, cfg, objects, None, True, False, True)
        _log("c2", t0, st[0], T)
    _fields("mid", st[1], out)
    while st[0] < T:
        st = yee.forward(st, cfg, objects, None, True, False, True)
        _log("c2", t0, st[0], T)
    _fields("fwd", st[1], out)
    for n in ("det_source", "det_thru", "det_cross"):
        out[n] = st[1].detector_states[n]["phasor"]
    tr = c2_transmissions(objects, st[1].detector_states)
    out["T_thru"], out["T_cross"] = tr["det_thru"], tr["det_cross"]
    print("transmissions", tr, flush=True)
    return out


if __name__ == "__main__":
    name = sys.argv[1]
    t0 = time.time()
    res = {"c1": make_c1, "c3a": make_c3a, "c3b": make_c3b, "c4": make_c4, "c2": make_c2}[name]()
    path = os.path.join(HERE, f"cfg_{name}.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in res.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB) in {time.time() - t0:.0f}s")
