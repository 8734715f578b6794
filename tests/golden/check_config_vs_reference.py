"""Config-faithful runs of the REFERENCE'S OWN SOURCE (oracle/refexec.py) against the oracle's config goldens.

    python tests/golden/check_config_vs_reference.py c3a | c1 | c2 | c3b | c4     (this container only; minutes)

Steps a BASELINE config at its real grid with the reference's ``update_E`` / ``update_H`` / ``update_detector_states``
executed from /root/reference under the NumPy ``jax.numpy`` stand-in, and compares with ``tests/golden/cfg_*.npz``
(oracle outputs, which the CUDA path is tested against in tests/test_config_parity.py).  Results are appended to
profiles/r02_config_vs_reference_source.txt.  Not part of the test suite (C1 takes ~15 min on the CPU)."""

from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import configs  # noqa: E402
from configs import field_norms, sub  # noqa: E402
from oracle import refexec  # noqa: E402
from scenes import rel_l2  # noqa: E402


def step(ref, robj, cfg, a, t):
    tt = ref.jnp.asarray(t, dtype=np.int32)
    H_prev = a.fields.H
    a = ref.update_E(tt, a, robj, cfg, True)
    a = ref.update_H(tt, a, robj, cfg, True)
    return ref.update_detector_states(tt, a, robj, cfg, H_prev, False)


def report(lines):
    path = os.path.join(ROOT, "profiles", "r02_config_vs_reference_source.txt")
    with open(path, "a") as f:
        for ln in lines:
            print(ln, flush=True)
            f.write(ln + "\n")


def main():
    which = sys.argv[1]
    ref = refexec.Reference()
    g = np.load(os.path.join(HERE, f"cfg_{which}.npz"))
    objects, arrays, cfg = getattr(configs, f"build_{which}")()
    T = cfg.time_steps_total
    robj = ref.wrap_objects(objects, cfg)
    a = refexec.to_jarr(arrays.reset())
    t0 = time.time()
    lines = [f"# {which}: reference source (fdtd/update.py etc. executed from /root/reference) vs oracle golden tests/golden/cfg_{which}.npz, grid {objects.volume.grid_shape}, {T} steps"]
    if which == "c3a":
        from make_config_golden import C3A_MID

        save = {}

        for t in range(T):
            a = step(ref, robj, cfg, a, t)
            if t + 1 == C3A_MID:
                for k, v in (("mid_E", a.fields.E), ("mid_H", a.fields.H), ("mid_P", a.fields.dispersive_P_curr)):
                    save[k] = np.asarray(v).copy()
                    lines.append(f"[C3a] {k} at step {C3A_MID}: rel-L2 {rel_l2(np.asarray(v), g[k]):.3e}")
        for n in ("pulse_trace_A", "pulse_trace_B"):
            lines.append(f"[C3a] {n} (all {T} steps): rel-L2 {rel_l2(np.asarray(a.detector_states[n]['fields']), g[n]):.3e}")
        np.savez_compressed(os.path.join(HERE, "cfg_c3a_refsrc.npz"), **save, fwd_E=np.asarray(a.fields.E), fwd_H=np.asarray(a.fields.H),
                            pulse_trace_A=np.asarray(a.detector_states["pulse_trace_A"]["fields"]), pulse_trace_B=np.asarray(a.detector_states["pulse_trace_B"]["fields"]))
        scale = float(np.linalg.norm(g["mid_E"].astype(np.float64)))
        lines.append(f"[C3a] final E: |ref - oracle| / |E(mid)| = {float(np.linalg.norm(np.asarray(a.fields.E, np.float64) - g['fwd_E'])) / scale:.3e}")
    elif which == "c2":
        # the bench scene at cells-per-lambda 6: the first C2_MID steps (pulse inside the coupling section)
        from make_config_golden import C2_MID

        for t in range(C2_MID):
            a = step(ref, robj, cfg, a, t)
            if t % 100 == 0:
                print(f"step {t}/{C2_MID} {time.time() - t0:.0f}s", flush=True)
        E, H = np.asarray(a.fields.E), np.asarray(a.fields.H)
        np.savez_compressed(os.path.join(HERE, "cfg_c2_refsrc.npz"), mid_E=sub(E), mid_H=sub(H), mid_E_norm=field_norms(E), mid_H_norm=field_norms(H))
        lines.append(f"[C2] E at step {C2_MID} (every 4th cell): rel-L2 {rel_l2(sub(E), g['mid_E']):.3e}")
        lines.append(f"[C2] H at step {C2_MID} (every 4th cell): rel-L2 {rel_l2(sub(H), g['mid_H']):.3e}")
        lines.append(f"[C2] |E| per component at step {C2_MID}: rel-L2 {rel_l2(field_norms(E), g['mid_E_norm']):.3e}")
    elif which in ("c1", "c4", "c3b"):
        for t in range(T):
            a = step(ref, robj, cfg, a, t)
            if t % 25 == 0:
                print(f"step {t}/{T} {time.time() - t0:.0f}s", flush=True)
        E, H = np.asarray(a.fields.E), np.asarray(a.fields.H)
        extra = {}
        if which == "c1":
            from configs import t_pick

            for key in ("XY Plane", "XZ Plane", "YZ Plane"):
                v = np.asarray(a.detector_states["Energy Video"][key])[:, ::2, ::2]
                extra[f"video_{key[:2]}"] = np.ascontiguousarray(v[t_pick(v.shape[0])])
                extra[f"video_{key[:2]}_norm"] = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2)))  # of the sub-sampled frames
        if which == "c4":
            for n, k in (("out flux", "out_flux"), ("in flux", "in_flux")):
                extra[k] = np.asarray(a.detector_states[n]["poynting_flux"])
        if which == "c3b":
            v = np.asarray(a.detector_states["Electric Field Video"]["fields"])
            extra["video_norm"] = np.sqrt((v.astype(np.float64).reshape(v.shape[0], -1) ** 2).sum(axis=1))
        np.savez_compressed(os.path.join(HERE, f"cfg_{which}_refsrc.npz"), fwd_E=sub(E), fwd_H=sub(H), fwd_E_norm=field_norms(E), fwd_H_norm=field_norms(H), **extra)
        lines.append(f"[{which.upper()}] fwd E (every 4th cell): rel-L2 {rel_l2(sub(E), g['fwd_E']):.3e}")
        lines.append(f"[{which.upper()}] fwd H (every 4th cell): rel-L2 {rel_l2(sub(H), g['fwd_H']):.3e}")
        lines.append(f"[{which.upper()}] fwd |E| per component: rel-L2 {rel_l2(field_norms(E), g['fwd_E_norm']):.3e}")
        if which == "c1":
            st = a.detector_states["Energy Video"]
            for key in ("XY Plane", "XZ Plane", "YZ Plane"):
                v = np.asarray(st[key])
                lines.append(f"[C1] video {key} norm of every frame: rel-L2 {rel_l2(np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2))), g[f'video_{key[:2]}_norm']):.3e}")
        elif which == "c4":
            for n, k in (("out flux", "out_flux"), ("in flux", "in_flux")):
                lines.append(f"[C4] {n} series: rel-L2 {rel_l2(extra[k], g[k]):.3e}")
        else:
            lines.append(f"[C3b] Ey video norm of every frame: rel-L2 {rel_l2(extra['video_norm'][1:], g['video_norm'][1:]):.3e}")
    lines.append(f"# ({time.time() - t0:.0f} s on the CPU)")
    report(lines)


if __name__ == "__main__":
    main()
