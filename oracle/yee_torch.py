"""Differentiable CPU oracle: the forward Yee step of ``oracle/yee.py`` restated with torch ops so
that ``torch.autograd`` provides the reference gradient (what ``jax.vjp(forward...)`` provides in
``fdtd/fdtd.py:227-250`` and what the checkpointed method differentiates, ``fdtd.py:482-493``).

TEST INFRASTRUCTURE ONLY (see ``oracle/yee.py``).  Scope: isotropic / diagonal materials with
optional conductivity, CPML, PEC/PMC, uniform or stretched grids, TFSF/dipole sources (constants
under ``stop_gradient``, ``tfsf.py:286-304``), field / energy / Poynting / phasor detectors with
exact interpolation.  Runs in float32 (like the product) or float64 (gradient reference).
"""

from __future__ import annotations

import numpy as np
import torch

from fdtdx_b200.boundaries import PerfectElectricConductor, PerfectMagneticConductor
from fdtdx_b200.constants import c as c0
from fdtdx_b200.constants import eta0
from fdtdx_b200.detectors import COMPONENT_NAMES, EnergyDetector, FieldDetector, PhasorDetector, PoyntingFluxDetector
from oracle import yee


def _pad(f, wrap):
    for ax in range(3):
        d = ax + 1
        if wrap[ax]:
            f = torch.cat([f.narrow(d, f.shape[d] - 1, 1), f, f.narrow(d, 0, 1)], dim=d)
        else:
            z = torch.zeros_like(f.narrow(d, 0, 1))
            f = torch.cat([z, f, z], dim=d)
    return f


def _scale(config, axis, stencil, dtype):
    s = yee._metric_scale(config, axis, None if not config.has_nonuniform_grid else config.resolved_grid.shape, stencil)
    return torch.as_tensor(np.asarray(s, dtype=np.float64), dtype=dtype)


def _curl(config, Fp, psi, objects, is_curl_E, dtype):
    """curl_E / curl_H with CPML (curl.py:227-397); psi' is returned but the caller may ignore it."""
    st = "forward" if is_curl_E else "backward"
    sx, sy, sz = (_scale(config, a, st, dtype) for a in range(3))
    Fx, Fy, Fz = Fp[0], Fp[1], Fp[2]
    c = (slice(1, -1),) * 3
    if is_curl_E:
        dyFz = (Fz[1:-1, 2:, 1:-1] - Fz[c]) * sy
        dzFy = (Fy[1:-1, 1:-1, 2:] - Fy[c]) * sz
        dzFx = (Fx[1:-1, 1:-1, 2:] - Fx[c]) * sz
        dxFz = (Fz[2:, 1:-1, 1:-1] - Fz[c]) * sx
        dxFy = (Fy[2:, 1:-1, 1:-1] - Fy[c]) * sx
        dyFx = (Fx[1:-1, 2:, 1:-1] - Fx[c]) * sy
    else:
        dyFz = (Fz[c] - Fz[1:-1, :-2, 1:-1]) * sy
        dzFy = (Fy[c] - Fy[1:-1, 1:-1, :-2]) * sz
        dzFx = (Fx[c] - Fx[1:-1, 1:-1, :-2]) * sz
        dxFz = (Fz[c] - Fz[:-2, 1:-1, 1:-1]) * sx
        dxFy = (Fy[c] - Fy[:-2, 1:-1, 1:-1]) * sx
        dyFx = (Fx[c] - Fx[1:-1, :-2, 1:-1]) * sy
    comps = [dyFz - dzFy, dzFx - dxFz, dxFy - dyFx]
    new_psi = {}
    for pml in objects.pml_objects:
        a = pml.axis
        i, j = (a + 1) % 3, (a + 2) % 3
        dj, di = {0: (dxFz, dxFy), 1: (dyFx, dyFz), 2: (dzFy, dzFx)}[a]
        gs = pml.grid_slice
        if is_curl_E:
            ca, cb, ik = pml.pml_a_H, pml.pml_b_H, pml.inv_kappa_H
        else:
            ca, cb, ik = pml.pml_a_E, pml.pml_b_E, pml.inv_kappa_E
        ca, cb, ik = (torch.as_tensor(np.asarray(v, np.float64), dtype=dtype) for v in (ca, cb, ik))
        p1, p2 = psi[pml.name]
        p1n = cb * p1 + ca * dj[gs]
        p2n = cb * p2 + ca * di[gs]
        if pml.kappa_is_one:
            c1, c2 = p1n, p2n
        else:
            c1 = (ik - 1.0) * dj[gs] + p1n
            c2 = (ik - 1.0) * di[gs] + p2n
        ci = comps[i].clone()
        cj = comps[j].clone()
        ci[gs] = ci[gs] - c1
        cj[gs] = cj[gs] + c2
        comps[i], comps[j] = ci, cj
        new_psi[pml.name] = (p1n, p2n)
    return torch.stack(comps), new_psi


def _source_delta(objects, arrays_np, config, t, which, shape, dtype):
    """Injected increments as constants: difference of the NumPy oracle's source loop on zeros."""
    zero = np.zeros((3, *shape), np.float32)
    out = yee._apply_sources(zero, arrays_np, objects, config, t, which, inverse=False)
    return torch.as_tensor(out.astype(np.float64), dtype=dtype)


def _mask(objects, kind, shape, dtype):
    m = torch.ones((3, *shape), dtype=dtype)
    for b in objects.boundary_objects:
        if isinstance(b, kind):
            for comp in b.tangential_components:
                m[(comp, *b.grid_slice)] = 0
    return m


def _bea(cur, prev, config, axis, dtype):
    if config is None or not config.has_nonuniform_grid:
        return 0.5 * (cur + prev)
    w = config.resolved_grid.cell_widths(axis).astype(np.float64)
    pw = np.concatenate([w[:1], w[:-1]])
    bs = [1, 1, 1]
    bs[axis] = cur.shape[axis]
    chw = torch.as_tensor(0.5 * w, dtype=dtype).reshape(bs)
    phw = torch.as_tensor(0.5 * pw, dtype=dtype).reshape(bs)
    return (cur * phw + prev * chw) / (chw + phw)


def interpolate_fields(Ep, Hp, config, dtype):
    Ex, Ey, Ez = Ep[0], Ep[1], Ep[2]
    Hx, Hy, Hz = Hp[0], Hp[1], Hp[2]
    b = lambda cur, prev, ax: _bea(cur, prev, config, ax, dtype)
    Exi = (b(Ex[1:-1, 1:-1, 1:-1], Ex[:-2, 1:-1, 1:-1], 0) + b(Ex[1:-1, 1:-1, 2:], Ex[:-2, 1:-1, 2:], 0)) / 2.0
    Eyi = (b(Ey[1:-1, 1:-1, 1:-1], Ey[1:-1, :-2, 1:-1], 1) + b(Ey[1:-1, 1:-1, 2:], Ey[1:-1, :-2, 2:], 1)) / 2.0
    Ezi = Ez[1:-1, 1:-1, 1:-1]
    Hxi = b(Hx[1:-1, 1:-1, 1:-1], Hx[1:-1, :-2, 1:-1], 1)
    Hyi = b(Hy[1:-1, 1:-1, 1:-1], Hy[:-2, 1:-1, 1:-1], 0)
    lo = b(b(Hz[1:-1, 1:-1, 1:-1], Hz[:-2, 1:-1, 1:-1], 0), b(Hz[1:-1, :-2, 1:-1], Hz[:-2, :-2, 1:-1], 0), 1)
    hi = b(b(Hz[1:-1, 1:-1, 2:], Hz[:-2, 1:-1, 2:], 0), b(Hz[1:-1, :-2, 2:], Hz[:-2, :-2, 2:], 0), 1)
    return torch.stack([Exi, Eyi, Ezi]), torch.stack([Hxi, Hyi, (lo + hi) / 2.0])


def _detector_update(det, t, E, H, state, inv_eps, inv_mu, config, dtype):
    sel = lambda: torch.stack([a for n, a in zip(COMPONENT_NAMES, (E[0], E[1], E[2], H[0], H[1], H[2])) if n in det.components])
    if isinstance(det, PhasorDetector):
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        ang = torch.as_tensor(det._angular_frequencies.astype(np.float64) * (t * config.time_step_duration), dtype=dtype)
        ph = torch.complex(torch.cos(ang), torch.sin(ang)).reshape(-1, 1, 1, 1, 1)
        new = sel()[None].to(cdt) * ph * det._static_scale() * float(det._window_at_time_step_arr[t])
        if det.reduce_volume:
            w = torch.as_tensor(det._cached_cell_volume_weights.astype(np.float64), dtype=dtype)
            new = (new * w[None, None]).sum(dim=(2, 3, 4)) / w.sum()
        return {"phasor": state["phasor"] - new[None] if det.inverse else state["phasor"] + new[None]}
    idx = int(det._time_step_to_arr_idx[t])
    if isinstance(det, EnergyDetector):
        e = 0.5 * ((E * E) / inv_eps).sum(0) if inv_eps.shape[0] == 3 else 0.5 * ((E * E) / inv_eps).sum(0)
        mu = inv_mu if torch.is_tensor(inv_mu) else torch.tensor(float(inv_mu), dtype=dtype)
        e = e + 0.5 * ((H * H) / mu).sum(0)
        out = {}
        if det.as_slices:
            if det.use_mean:
                vals = {"XY Plane": e.mean(2), "XZ Plane": e.mean(1), "YZ Plane": e.mean(0)}
            else:
                xi, yi, zi = det._slice_indices
                vals = {"XY Plane": e[:, :, zi], "XZ Plane": e[:, yi, :], "YZ Plane": e[xi]}
        elif det.reduce_volume:
            w = torch.as_tensor(det._cached_cell_volume_weights.astype(np.float64), dtype=dtype)
            vals = {"energy": (e * w).sum().reshape(1)}
        else:
            vals = {"energy": e}
    elif isinstance(det, FieldDetector):
        EH = sel()
        if det.reduce_volume:
            w = torch.as_tensor(det._cached_cell_volume_weights.astype(np.float64), dtype=dtype)
            EH = (EH * w[None]).sum(dim=(1, 2, 3)) / w.sum()
        vals = {"fields": EH}
    elif isinstance(det, PoyntingFluxDetector):
        pf = torch.stack([E[1] * H[2] - E[2] * H[1], E[2] * H[0] - E[0] * H[2], E[0] * H[1] - E[1] * H[0]])
        if not det.keep_all_components:
            pf = pf[det.propagation_axis]
        if det.direction == "-":
            pf = -pf
        if det.reduce_volume:
            w = torch.as_tensor(np.asarray(det._cached_face_area_weights, np.float64), dtype=dtype)
            pf = pf * w
            pf = pf.sum(dim=(1, 2, 3)) if det.keep_all_components else pf.sum().reshape(1)
        vals = {"poynting_flux": pf}
    else:
        raise NotImplementedError(type(det))
    out = {}
    for k, v in vals.items():
        s = state[k]
        out[k] = torch.cat([s[:idx], v[None], s[idx + 1 :]], dim=0)
    return out


class Stepper:
    """Holds the per-run constants; ``step`` is one differentiable forward step (forward.py:83-156)."""

    def __init__(self, arrays_np, objects, config, dtype=torch.float64):
        self.arrays_np, self.objects, self.config, self.dtype = arrays_np, objects, config, dtype
        self.shape = objects.volume.grid_shape
        T = self.T
        self.sE = None if arrays_np.electric_conductivity is None else T(arrays_np.electric_conductivity)
        self.sH = None if arrays_np.magnetic_conductivity is None else T(arrays_np.magnetic_conductivity)
        self.wrap = yee.get_wrap_padding_axes(objects)
        self.mE = _mask(objects, PerfectElectricConductor, self.shape, dtype)
        self.mH = _mask(objects, PerfectMagneticConductor, self.shape, dtype)

    def T(self, a):
        return torch.as_tensor(np.asarray(a, np.float64), dtype=self.dtype)

    def initial(self, arrays_np=None):
        a = arrays_np or self.arrays_np
        cdt = torch.complex64 if self.dtype == torch.float32 else torch.complex128
        psiE = {k: (self.T(x), self.T(y)) for k, (x, y) in a.fields.psi_E.items()}
        psiH = {k: (self.T(x), self.T(y)) for k, (x, y) in a.fields.psi_H.items()}
        det = {k: {k2: torch.as_tensor(v2, dtype=cdt if np.iscomplexobj(v2) else self.dtype) for k2, v2 in v.items()} for k, v in a.detector_states.items()}
        return self.T(a.fields.E), self.T(a.fields.H), psiE, psiH, det

    def step(self, t, E, H, psiE, psiH, det, inv_eps, inv_mu, ade=None):
        """``ade``: None or a dict with P, Pprev (n_poles, 3, *shape) and c1, c2, c3, c4 (n_poles, 1|3, *shape;
        c4 may be None); the updated P / Pprev are written back into the dict (update.py:316-350)."""
        config, objects, dtype, shape = self.config, self.objects, self.dtype, self.shape
        c = config.courant_number
        H_prev = H
        K, psiE = _curl(config, _pad(H, self.wrap), psiE, objects, False, dtype)
        s = c * self.sE * eta0 * inv_eps / 2 if self.sE is not None else None
        if ade is not None:
            P, Q = ade["P"], ade["Pprev"]
            P_hat = ade["c1"] * P + ade["c2"] * Q + ade["c3"] * E
            E = (E if s is None else (1 - s) * E) + c * K * inv_eps + inv_eps * (P - P_hat).sum(0)
            den = 1.0 if s is None else 1 + s
            if ade.get("c4") is not None:
                den = den + inv_eps * ade["c4"].sum(0)
            E = E / den
            ade["Pprev"] = P
            ade["P"] = P_hat + ade["c4"] * E if ade.get("c4") is not None else P_hat
        elif s is not None:
            E = ((1 - s) * E + c * K * inv_eps) / (1 + s)
        else:
            E = E + c * K * inv_eps
        E = (E + _source_delta(objects, self.arrays_np, config, t, "E", shape, dtype)) * self.mE
        K, psiH = _curl(config, _pad(E, self.wrap), psiH, objects, True, dtype)
        if self.sH is not None:
            s = c * self.sH / eta0 * inv_mu / 2
            H = ((1 - s) * H - c * K * inv_mu) / (1 + s)
        else:
            H = H - c * K * inv_mu
        H = (H + _source_delta(objects, self.arrays_np, config, t, "H", shape, dtype)) * self.mH
        on = [d for d in objects.forward_detectors if bool(d._is_on_at_time_step_arr[t])]
        det = dict(det)
        if on:
            Ei, Hi = interpolate_fields(_pad(E, self.wrap), _pad((H_prev + H) / 2, self.wrap), config, dtype)
            for d in on:
                gs = d.grid_slice
                if d.exact_interpolation:
                    Er, Hr = Ei[(slice(None), *gs)], Hi[(slice(None), *gs)]
                else:
                    Er, Hr = E[(slice(None), *gs)], H[(slice(None), *gs)]
                mu_r = inv_mu[(slice(None), *gs)] if torch.is_tensor(inv_mu) else inv_mu
                det[d.name] = _detector_update(d, t, Er, Hr, det[d.name], inv_eps[(slice(None), *gs)], mu_r, config, dtype)
        return E, H, psiE, psiH, det


def run_forward(arrays_np, objects, config, steps, inv_eps=None, inv_mu=None, dtype=torch.float64, E0=None, H0=None, coeffs=None):
    """Runs ``steps`` forward steps from the container's state (diagonal tier); ``inv_eps`` /
    ``inv_mu`` (and, for dispersive media, ``coeffs = {"c1": ..., "c2": ..., "c3": ..., "c4": ...}``) may be
    leaf tensors requiring grad.  Returns (E, H, detector_states)."""
    S = Stepper(arrays_np, objects, config, dtype)
    E, H, psiE, psiH, det = S.initial()
    if E0 is not None:
        E = E0
    if H0 is not None:
        H = H0
    if inv_eps is None:
        inv_eps = S.T(arrays_np.inv_permittivities)
    mu_np = arrays_np.inv_permeabilities
    if inv_mu is None:
        inv_mu = S.T(mu_np) if isinstance(mu_np, np.ndarray) else float(mu_np)
    ade = None
    if arrays_np.dispersive_c1 is not None:
        get = lambda k: (coeffs[k] if coeffs is not None and coeffs.get(k) is not None else (None if getattr(arrays_np, "dispersive_" + k) is None else S.T(getattr(arrays_np, "dispersive_" + k))))
        ade = {"P": S.T(arrays_np.fields.dispersive_P_curr), "Pprev": S.T(arrays_np.fields.dispersive_P_prev), "c1": get("c1"), "c2": get("c2"), "c3": get("c3"), "c4": get("c4")}
    for t in range(steps):
        E, H, psiE, psiH, det = S.step(t, E, H, psiE, psiH, det, inv_eps, inv_mu, ade)
    return E, H, det


def reversible_gradient(arrays_final_np, objects, config, loss_fn, dtype=torch.float64, progress=None):
    """The reference's ``fdtd_bwd`` (fdtd/fdtd.py:262-333) restated: starting from the final state,
    repeat { ``backward`` (NumPy oracle, reset_fields=False) -> VJP of one forward step at the
    reconstructed state with the frozen final psi }, accumulating d loss / d inv_eps (and inv_mu).
    ``loss_fn(E, H, det)`` defines the output cotangents.  The extra t = -1 iteration of the
    reference (starting from the all-zero state 0) is omitted, as in the product."""
    T_total = config.time_steps_total
    S = Stepper(arrays_final_np, objects, config, dtype)
    E, H, psiE, psiH, det = S.initial(arrays_final_np)
    leaves = [E, H] + [v for st in det.values() for v in st.values()]
    for x in leaves:
        x.requires_grad_(True)
    loss = loss_fn(E, H, det)
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    zero = lambda x, g: torch.zeros_like(x) if g is None else g
    lamE, lamH = zero(E, grads[0]), zero(H, grads[1])
    names = [(d, k) for d, st in det.items() for k in st]
    lamdet = {n: zero(det[n[0]][n[1]], g) for n, g in zip(names, grads[2:])}
    lampsiE = {k: (torch.zeros_like(a), torch.zeros_like(b)) for k, (a, b) in psiE.items()}
    lampsiH = {k: (torch.zeros_like(a), torch.zeros_like(b)) for k, (a, b) in psiH.items()}
    mu_np = arrays_final_np.inv_permeabilities
    g_eps = torch.zeros(arrays_final_np.inv_permittivities.shape, dtype=dtype)
    g_mu = torch.zeros(mu_np.shape, dtype=dtype) if isinstance(mu_np, np.ndarray) else None
    state = (T_total, arrays_final_np)
    for t in range(T_total - 1, -1, -1):
        if progress is not None:
            progress(t)
        state = yee.backward(state, config, objects, None, record_detectors=False, reset_fields=False)
        a = state[1]
        Et, Ht = S.T(a.fields.E).requires_grad_(True), S.T(a.fields.H).requires_grad_(True)
        pE = {k: (x.clone().requires_grad_(True), y.clone().requires_grad_(True)) for k, (x, y) in psiE.items()}
        pH = {k: (x.clone().requires_grad_(True), y.clone().requires_grad_(True)) for k, (x, y) in psiH.items()}
        ie = S.T(a.inv_permittivities).requires_grad_(True)
        im = S.T(mu_np).requires_grad_(True) if isinstance(mu_np, np.ndarray) else float(mu_np)
        d_in = {k: {k2: v2.detach().clone().requires_grad_(True) for k2, v2 in v.items()} for k, v in det.items()}
        E1, H1, pE1, pH1, d1 = S.step(t, Et, Ht, pE, pH, d_in, ie, im)
        outs = [E1, H1] + [x for k in pE1 for x in pE1[k]] + [x for k in pH1 for x in pH1[k]] + [d1[n[0]][n[1]] for n in names]
        cots = [lamE, lamH] + [x for k in pE1 for x in lampsiE[k]] + [x for k in pH1 for x in lampsiH[k]] + [lamdet[n] for n in names]
        ins = [Et, Ht] + [x for k in pE for x in pE[k]] + [x for k in pH for x in pH[k]] + [ie] + ([im] if g_mu is not None else [])
        keep = [(o, c) for o, c in zip(outs, cots) if o.requires_grad]
        gr = torch.autograd.grad([o for o, _ in keep], ins, grad_outputs=[c for _, c in keep], allow_unused=True)
        gi = iter(gr)
        lamE, lamH = zero(Et, next(gi)), zero(Ht, next(gi))
        lampsiE = {k: (zero(pE[k][0], next(gi)), zero(pE[k][1], next(gi))) for k in pE}
        lampsiH = {k: (zero(pH[k][0], next(gi)), zero(pH[k][1], next(gi))) for k in pH}
        g_eps = g_eps + zero(ie, next(gi))
        if g_mu is not None:
            g_mu = g_mu + zero(im, next(gi))
    return g_eps, g_mu
