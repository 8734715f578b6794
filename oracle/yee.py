"""CPU oracle: NumPy float32 restatement of fdtdx's Yee time-stepping hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fdtdx_b200/`` imports this module; it is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the checker / CPU baseline, never as a product path.

Parity status: the reference (pure Python on JAX, v0.6.2) cannot be imported in this image (no
jax / equinox / pytreeclass; no network) and ships no golden field vectors, so this oracle is
pinned by (i) following the reference files line by line in the reference's own op order and
float32 rounding sequence, and (ii) the reference's closed-form unit tests and physics tests
re-typed against it in ``tests/test_oracle_*.py``.  Bit-level parity with
a real fdtdx/JAX run is therefore argued, not measured: "parity unpinned against reference
outputs; pinned against the reference's known-answer tests".

Every function cites the reference lines it follows (paths relative to
``/root/reference/src/fdtdx/``).
"""

from __future__ import annotations

from dataclasses import replace

import numpy as np

from fdtdx_b200.boundaries import (
    BlochBoundary,
    PerfectElectricConductor,
    PerfectlyMatchedLayer,
    PerfectMagneticConductor,
)
from fdtdx_b200.constants import c as c0
from fdtdx_b200.constants import eta0
from fdtdx_b200.container import ArrayContainer, RecordingState, _TorchLeaf
from fdtdx_b200.detectors import (
    COMPONENT_NAMES,
    ClosedSurfacePhasorPoyntingFluxDetector,
    ClosedSurfacePoyntingFluxDetector,
    EnergyDetector,
    FieldDetector,
    PhasorDetector,
    PoyntingFluxDetector,
)
from fdtdx_b200.recorder import REC_F32
from fdtdx_b200.sources import PointDipoleSource, TFSFPlaneSource, get_oriented_transverse_axes

F = np.float32


# ----------------------------------------------------------------------------------------------
# padding  (core/misc.py:615-641, fdtd/update.py:29-45, 92-136)
# ----------------------------------------------------------------------------------------------
def get_wrap_padding_axes(objects) -> tuple[bool, bool, bool]:
    wrap = [False, False, False]
    for b in objects.boundary_objects:
        if b.uses_wrap_padding:
            wrap[b.axis] = True
    return tuple(wrap)


def _C(x):
    """Cast to the field dtype: float32, or complex64 for the complex fields of a Bloch run with
    k != 0 (initialization.py:581-596)."""
    x = np.asarray(x)
    return x.astype(np.complex64) if np.iscomplexobj(x) else x.astype(F)


def pad_fields(fields: np.ndarray, periodic_axes) -> np.ndarray:
    padded = fields
    for i, periodic in enumerate(periodic_axes):
        pw = [(0, 0)] * 4
        pw[i + 1] = (1, 1)
        padded = np.pad(padded, pw, mode="wrap" if periodic else "constant")
    return padded


def pad_fields_for_boundaries(fields, objects, config) -> np.ndarray:
    periodic_axes = get_wrap_padding_axes(objects)
    padded = pad_fields(fields, periodic_axes)
    for axis in range(3):
        if config.symmetry[axis] != 0 and periodic_axes[axis]:
            idx = [slice(None)] * 4
            idx[axis + 1] = slice(0, 1)
            padded[tuple(idx)] = 0
    # BlochBoundary.apply_pad_correction (bloch.py:61-96): the ghost plane that wrapped from the far
    # side is multiplied by conj(phase) on the '-' face and by phase on the '+' face
    shape = objects.volume.grid_shape
    for b in objects.boundary_objects:
        if isinstance(b, BlochBoundary) and b.needs_complex_fields:
            phase = b.get_bloch_phase(shape, config)
            idx = [slice(None)] * 4
            if b.direction == "-":
                idx[b.axis + 1] = 0
                padded[tuple(idx)] = (padded[tuple(idx)] * np.conj(phase)).astype(np.complex64)
            else:
                idx[b.axis + 1] = -1
                padded[tuple(idx)] = (padded[tuple(idx)] * phase).astype(np.complex64)
    return padded


def field_component_parity(field_type: str, component: int, axis: int, wall: int) -> int:
    """core/physics/symmetry.py: mirror parity of a component across a plane normal to ``axis``
    (wall -1: PEC / electric, +1: PMC / magnetic)."""
    normal = component == axis
    if wall == -1:
        return (1 if normal else -1) if field_type == "E" else (-1 if normal else 1)
    if wall == 1:
        return (-1 if normal else 1) if field_type == "E" else (1 if normal else -1)
    raise ValueError(f"wall must be -1 (PEC) or +1 (PMC), got {wall}")


def mirror_pairs_on_plane(field_type: str, component: int, axis: int, wall: int) -> bool:
    """core/physics/symmetry.py: an electric plane sits on the tangential-E node row, so components
    sampled there (E tangential, H normal) pair as m +- j; everything else is a plain flip."""
    sits = (component != axis) if field_type == "E" else (component == axis)
    return wall == -1 and sits


def pad_fields_with_symmetry_mirror(fields, objects, config, field_type: str) -> np.ndarray:
    """update.py:139-198: the halo of every *electric* ``config.symmetry`` plane holds the parity-weighted
    mirror partner (detector co-location stencil only; the field updates never see it)."""
    padded = pad_fields_for_boundaries(fields, objects, config)
    for b in objects.boundary_objects:
        if not getattr(b, "_is_symmetry_wall", False):
            continue
        axis = b.axis
        wall = config.symmetry[axis]
        if wall != -1:
            continue
        for component in range(3):
            parity = field_component_parity(field_type, component, axis, wall)
            src = 2 if mirror_pairs_on_plane(field_type, component, axis, wall) else 1
            tgt_i = [slice(None)] * 3
            src_i = [slice(None)] * 3
            tgt_i[axis] = slice(0, 1)
            src_i[axis] = slice(src, src + 1)
            padded[(component, *tgt_i)] = parity * padded[(component, *src_i)]
    return padded


# ----------------------------------------------------------------------------------------------
# curl + CPML  (core/physics/curl.py:10-39, 227-397; perfectly_matched_layer.py:138-190)
# ----------------------------------------------------------------------------------------------
def _metric_scale(config, axis: int, shape, stencil: str):
    if not config.has_nonuniform_grid:
        return F(1.0)
    grid = config.resolved_grid
    widths = grid.cell_widths(axis)
    if stencil == "backward":
        prev_widths = np.concatenate([widths[:1], widths[:-1]])
        widths = F(0.5) * (widths + prev_widths)
    elif stencil != "forward":
        raise ValueError(f"Unknown derivative stencil: {stencil}")
    reference_spacing = c0 * config.time_step_duration / config.courant_number
    scale = (F(reference_spacing) / widths).astype(F)
    bs = [1, 1, 1]
    bs[axis] = shape[axis]
    return scale.reshape(bs)


def step_cpml(pml: PerfectlyMatchedLayer, d1, d2, psi_1, psi_2, is_curl_E: bool, simulate_boundaries: bool):
    if is_curl_E:
        a, b, inv_kappa = pml.pml_a_H, pml.pml_b_H, pml.inv_kappa_H
    else:
        a, b, inv_kappa = pml.pml_a_E, pml.pml_b_E, pml.inv_kappa_E
    if simulate_boundaries:
        psi_1_new = b * psi_1 + a * d1
        psi_2_new = b * psi_2 + a * d2
    else:
        psi_1_new, psi_2_new = psi_1, psi_2
    if pml.kappa_start == 1.0 and pml.kappa_end == 1.0:
        corr_1, corr_2 = psi_1_new, psi_2_new
    else:
        corr_1 = (inv_kappa - F(1.0)) * d1 + psi_1_new
        corr_2 = (inv_kappa - F(1.0)) * d2 + psi_2_new
    return corr_1, corr_2, psi_1_new, psi_2_new


def curl_E(config, E_pad, psi_H, objects, simulate_boundaries: bool):
    shape = (E_pad.shape[1] - 2, E_pad.shape[2] - 2, E_pad.shape[3] - 2)
    sx = _metric_scale(config, 0, shape, "forward")
    sy = _metric_scale(config, 1, shape, "forward")
    sz = _metric_scale(config, 2, shape, "forward")
    Ex, Ey, Ez = E_pad[0], E_pad[1], E_pad[2]
    c = (slice(1, -1), slice(1, -1), slice(1, -1))
    dyEz = (Ez[1:-1, 2:, 1:-1] - Ez[c]) * sy
    dzEy = (Ey[1:-1, 1:-1, 2:] - Ey[c]) * sz
    dzEx = (Ex[1:-1, 1:-1, 2:] - Ex[c]) * sz
    dxEz = (Ez[2:, 1:-1, 1:-1] - Ez[c]) * sx
    dxEy = (Ey[2:, 1:-1, 1:-1] - Ey[c]) * sx
    dyEx = (Ex[1:-1, 2:, 1:-1] - Ex[c]) * sy
    comps = [dyEz - dzEy, dzEx - dxEz, dxEy - dyEx]
    psi_new = {}
    for pml in objects.pml_objects:
        a = pml.axis
        i, j = (a + 1) % 3, (a + 2) % 3
        if a == 0:
            dj, di = dxEz, dxEy
        elif a == 1:
            dj, di = dyEx, dyEz
        else:
            dj, di = dzEy, dzEx
        gs = pml.grid_slice
        p1, p2 = psi_H[pml.name]
        c1, c2, p1n, p2n = step_cpml(pml, dj[gs], di[gs], p1, p2, True, simulate_boundaries)
        comps[i] = comps[i].copy()
        comps[j] = comps[j].copy()
        comps[i][gs] += -c1
        comps[j][gs] += c2
        psi_new[pml.name] = (p1n, p2n)
    return _C(np.stack(comps, axis=0)), psi_new


def curl_H(config, H_pad, psi_E, objects, simulate_boundaries: bool):
    shape = (H_pad.shape[1] - 2, H_pad.shape[2] - 2, H_pad.shape[3] - 2)
    sx = _metric_scale(config, 0, shape, "backward")
    sy = _metric_scale(config, 1, shape, "backward")
    sz = _metric_scale(config, 2, shape, "backward")
    Hx, Hy, Hz = H_pad[0], H_pad[1], H_pad[2]
    c = (slice(1, -1), slice(1, -1), slice(1, -1))
    dyHz = (Hz[c] - Hz[1:-1, :-2, 1:-1]) * sy
    dzHy = (Hy[c] - Hy[1:-1, 1:-1, :-2]) * sz
    dzHx = (Hx[c] - Hx[1:-1, 1:-1, :-2]) * sz
    dxHz = (Hz[c] - Hz[:-2, 1:-1, 1:-1]) * sx
    dxHy = (Hy[c] - Hy[:-2, 1:-1, 1:-1]) * sx
    dyHx = (Hx[c] - Hx[1:-1, :-2, 1:-1]) * sy
    comps = [dyHz - dzHy, dzHx - dxHz, dxHy - dyHx]
    psi_new = {}
    for pml in objects.pml_objects:
        a = pml.axis
        i, j = (a + 1) % 3, (a + 2) % 3
        if a == 0:
            dj, di = dxHz, dxHy
        elif a == 1:
            dj, di = dyHx, dyHz
        else:
            dj, di = dzHy, dzHx
        gs = pml.grid_slice
        p1, p2 = psi_E[pml.name]
        c1, c2, p1n, p2n = step_cpml(pml, dj[gs], di[gs], p1, p2, False, simulate_boundaries)
        comps[i] = comps[i].copy()
        comps[j] = comps[j].copy()
        comps[i][gs] += -c1
        comps[j][gs] += c2
        psi_new[pml.name] = (p1n, p2n)
    return _C(np.stack(comps, axis=0)), psi_new


# ----------------------------------------------------------------------------------------------
# anisotropic helpers  (fdtd/misc.py:69-213, core/misc.py:564-612, fdtd/update.py:201-229)
# ----------------------------------------------------------------------------------------------
def expand_to_3x3(arr):
    if arr is None:
        return None
    arr = np.asarray(arr, dtype=F)
    if arr.ndim == 0:
        out = np.zeros((3, 3, 1, 1, 1), F)
        for i in range(3):
            out[i, i] = arr
        return out
    n, sp = arr.shape[0], arr.shape[1:]
    if n == 9:
        return arr.reshape((3, 3, *sp))
    out = np.zeros((3, 3, *sp), F)
    for i in range(3):
        out[i, i] = arr[0] if n == 1 else arr[i]
    return out


_AB_CACHE: dict = {}


def _update_matrices_cached(inv_raw, sigma_raw, c: float, eta_factor: float, reverse: bool = False):
    """The reference re-solves A, B every step (update.py:363); they depend on the materials only, so
    the result is memoised per (un-expanded) material array - value-identical, SURVEY App. C.5."""
    key = (id(inv_raw), None if sigma_raw is None else id(sigma_raw), float(c), float(eta_factor), bool(reverse))
    hit = _AB_CACHE.get(key)
    if hit is not None and hit[0] is inv_raw and hit[1] is sigma_raw:
        return hit[2], hit[3]
    A, B = compute_anisotropic_update_matrices(expand_to_3x3(inv_raw), expand_to_3x3(sigma_raw), c, eta_factor, reverse)
    if len(_AB_CACHE) > 8:
        _AB_CACHE.clear()
    _AB_CACHE[key] = (inv_raw, sigma_raw, A, B)
    return A, B


def compute_anisotropic_update_matrices(inv_prop, sigma, c: float, eta_factor: float, reverse: bool = False):
    sp = np.broadcast_shapes(inv_prop.shape[2:], (1, 1, 1) if sigma is None else sigma.shape[2:])
    eye = np.broadcast_to(np.eye(3, dtype=F)[:, :, None, None, None], (3, 3, *sp))
    M1, M2 = eye.copy(), eye.copy()
    inv_b = np.broadcast_to(inv_prop, (3, 3, *sp))
    if sigma is not None:
        factor = F(c * eta_factor / 2) * np.einsum("ijxyz,jkxyz->ikxyz", inv_b, np.broadcast_to(sigma, (3, 3, *sp))).astype(F)
        M1 = M1 + factor
        M2 = M2 - factor
    perm, inv_perm = (2, 3, 4, 0, 1), (3, 4, 0, 1, 2)
    if reverse:
        M1, M2 = M2, M1
    A = np.linalg.solve(M1.transpose(perm), M2.transpose(perm)).transpose(inv_perm).astype(F)
    B = (F(c) * np.linalg.solve(M1.transpose(perm), inv_b.transpose(perm)).transpose(inv_perm)).astype(F)
    return A, B


def get_anisotropic_averaging_widths(config):
    if not config.has_nonuniform_grid:
        return None
    grid = config.resolved_grid
    out = []
    for axis in range(3):
        w = grid.cell_widths(axis)
        padded = np.concatenate([w[:1], w, w[-1:]])
        bs = [1, 1, 1]
        bs[axis] = padded.shape[0]
        out.append(padded.reshape(bs))
    return tuple(out)


def avg_anisotropic_E_component(field, component: int, location: int, aniso_widths=None):
    s = field[component]
    if aniso_widths is None:
        return (
            (s + np.roll(s, -1, axis=location) + np.roll(s, 1, axis=component) + np.roll(s, (-1, 1), axis=(location, component)))
            / F(4)
        )[1:-1, 1:-1, 1:-1]
    centered = F(0.5) * (s + np.roll(s, -1, axis=location))
    width = aniso_widths[component]
    prev_w = np.roll(width, 1, axis=component)
    on_edge = (centered * prev_w + np.roll(centered, 1, axis=component) * width) / (width + prev_w)
    return on_edge[1:-1, 1:-1, 1:-1]


def avg_anisotropic_H_component(field, component: int, location: int, aniso_widths=None):
    s = field[component]
    if aniso_widths is None:
        return (
            (s + np.roll(s, 1, axis=location) + np.roll(s, -1, axis=component) + np.roll(s, (1, -1), axis=(location, component)))
            / F(4)
        )[1:-1, 1:-1, 1:-1]
    width = aniso_widths[location]
    prev_w = np.roll(width, 1, axis=location)
    on_edge = (s * prev_w + np.roll(s, 1, axis=location) * width) / (width + prev_w)
    centered = F(0.5) * (on_edge + np.roll(on_edge, -1, axis=component))
    return centered[1:-1, 1:-1, 1:-1]


def _tensor_apply(field, curl, A, B, objects, config, kind: str, sign: float):
    """Shared 3x3 apply of update.py:365-492 (E) / :761-822 (H): averages of the field and of the
    curl at each component's Yee location, then ``A@F + sign * B@K``."""
    avg = avg_anisotropic_E_component if kind == "E" else avg_anisotropic_H_component
    F_pad = pad_fields_for_boundaries(field, objects, config)
    K_pad = pad_fields_for_boundaries(curl, objects, config)
    w = get_anisotropic_averaging_widths(config)
    out = []
    for r in range(3):
        fa = [field[q] if q == r else avg(F_pad, component=q, location=r, aniso_widths=w) for q in range(3)]
        ka = [curl[q] if q == r else avg(K_pad, component=q, location=r, aniso_widths=w) for q in range(3)]
        t1 = A[r, 0] * fa[0] + A[r, 1] * fa[1] + A[r, 2] * fa[2]
        t2 = B[r, 0] * ka[0] + B[r, 1] * ka[1] + B[r, 2] * ka[2]
        out.append(t1 + t2 if sign > 0 else t1 - t2)
    return np.stack(out, axis=0).astype(F)


# ----------------------------------------------------------------------------------------------
# sources  (objects/sources/tfsf.py:193-409, 739-806; dipole.py:195-277)
# ----------------------------------------------------------------------------------------------
def _amplitude(src, time_step_f, offsets, config, quadrature: bool = False):
    """amp = profile((t + off) * dt) * static   (tfsf.py:250-258 / 351-359), float32; ``quadrature``:
    the carrier phase shifted by -pi/2 (tfsf.py:268-279)."""
    time = (F(time_step_f) + offsets).astype(F) * F(config.time_step_duration)
    ph = src.wave_character.phase_shift - (0.5 * np.pi if quadrature else 0.0)
    amp = src.temporal_profile.get_amplitude(time=time, period=src.wave_character.get_period(), phase_shift=ph)
    return (amp * F(src.static_amplitude_factor)).astype(F)


def tfsf_update_E(src: TFSFPlaneSource, E, inv_eps, time_step_f, inverse: bool, config):
    sign = 1 if src.direction == "+" else -1
    if inverse:
        sign = -sign
    n = src.propagation_axis
    a_ax, b_ax = get_oriented_transverse_axes(n)
    c = F(config.courant_number) * F(src.metric_scale_at_plane(config, "backward"))
    gs = src.grid_slice
    full = inv_eps.shape[0] == 9
    eps_sl = inv_eps[(slice(None), *gs)]
    amp = {}
    for ax in (a_ax, b_ax):
        if src._temporal_H_filter is None:
            amp[ax] = _amplitude(src, time_step_f, src._time_offset_H[ax], config)
        else:
            idx = (F(time_step_f) + src._time_offset_H[ax]).astype(F)
            filt = src._temporal_H_filter.astype(F)
            xp = np.arange(filt.shape[0], dtype=F)
            amp[ax] = (np.interp(idx, xp, filt, left=0.0, right=0.0).astype(F) * F(src.static_amplitude_factor)).astype(F)
    if np.iscomplexobj(src._H) and src._temporal_H_filter is None:
        # complex (lossy-mode) profile: Re * amp + Im * amp(phase - pi/2)   (tfsf.py:266-283)
        ampq = {ax: _amplitude(src, time_step_f, src._time_offset_H[ax], config, quadrature=True) for ax in (a_ax, b_ax)}
        Hb = (np.real(src._H[b_ax]).astype(F) * amp[b_ax] + np.imag(src._H[b_ax]).astype(F) * ampq[b_ax]).astype(F)
        Ha = (np.real(src._H[a_ax]).astype(F) * amp[a_ax] + np.imag(src._H[a_ax]).astype(F) * ampq[a_ax]).astype(F)
    else:
        Hb = (np.real(src._H[b_ax]).astype(F) * amp[b_ax]).astype(F)
        Ha = (np.real(src._H[a_ax]).astype(F) * amp[a_ax]).astype(F)
    E = E.copy()
    if full:
        for row in (n, a_ax, b_ax):
            corr = c * (eps_sl[row * 3 + a_ax] * (+Hb) + eps_sl[row * 3 + b_ax] * (-Ha))
            E[(row, *gs)] += F(sign) * corr
        return E
    ia = min(a_ax, eps_sl.shape[0] - 1)  # JAX clamps the static out-of-range index for 1-comp eps
    ib = min(b_ax, eps_sl.shape[0] - 1)
    Hb = Hb * c * eps_sl[ia]
    Ha = Ha * c * eps_sl[ib]
    E[(a_ax, *gs)] += F(sign) * Hb
    E[(b_ax, *gs)] += F(-sign) * Ha
    return E


def tfsf_update_H(src: TFSFPlaneSource, H, inv_mu, time_step_f, inverse: bool, config):
    sign = 1 if src.direction == "+" else -1
    if inverse:
        sign = -sign
    n = src.propagation_axis
    a_ax, b_ax = get_oriented_transverse_axes(n)
    c = F(config.courant_number) * F(src.metric_scale_at_plane(config, "forward"))
    gs = src.grid_slice
    is_arr = isinstance(inv_mu, np.ndarray) and inv_mu.ndim > 0
    full = is_arr and inv_mu.shape[0] == 9
    mu_sl = inv_mu[(slice(None), *gs)] if is_arr else inv_mu
    amp = {ax: _amplitude(src, time_step_f, src._time_offset_E[ax], config) for ax in (a_ax, b_ax)}
    if np.iscomplexobj(src._E):  # tfsf.py:366-383
        ampq = {ax: _amplitude(src, time_step_f, src._time_offset_E[ax], config, quadrature=True) for ax in (a_ax, b_ax)}
        Ea = (np.real(src._E[a_ax]).astype(F) * amp[a_ax] + np.imag(src._E[a_ax]).astype(F) * ampq[a_ax]).astype(F)
        Eb = (np.real(src._E[b_ax]).astype(F) * amp[b_ax] + np.imag(src._E[b_ax]).astype(F) * ampq[b_ax]).astype(F)
    else:
        Ea = (src._E[a_ax] * amp[a_ax]).astype(F)
        Eb = (src._E[b_ax] * amp[b_ax]).astype(F)
    H = H.copy()
    if full:
        for row in (n, a_ax, b_ax):
            corr = c * (mu_sl[row * 3 + a_ax] * (-Eb) + mu_sl[row * 3 + b_ax] * (+Ea))
            H[(row, *gs)] += F(sign) * corr
        return H
    if is_arr:
        ia = min(a_ax, mu_sl.shape[0] - 1)
        ib = min(b_ax, mu_sl.shape[0] - 1)
        Ea = Ea * c * mu_sl[ib]
        Eb = Eb * c * mu_sl[ia]
    else:
        Ea = Ea * c * F(mu_sl)
        Eb = Eb * c * F(mu_sl)
    H[(b_ax, *gs)] += F(sign) * Ea
    H[(a_ax, *gs)] += F(-sign) * Eb
    return H


def dipole_update(src: PointDipoleSource, field, inv_mat, time_step_f, inverse: bool, config, which: str):
    if (which == "E") != (src.source_type == "electric"):
        return field
    amp = src.temporal_profile.get_amplitude(
        time=(F(time_step_f) * F(config.time_step_duration)),
        period=src.wave_character.get_period(),
        phase_shift=src.wave_character.phase_shift,
    )
    sign = F(-1.0) if not inverse else F(1.0)
    gs = src.grid_slice
    is_arr = isinstance(inv_mat, np.ndarray) and inv_mat.ndim > 0
    # scale = c * amplitude * static * amp : python-double product times the float32 profile value
    scale = F(config.courant_number * src.amplitude * src.static_amplitude_factor) * F(amp)
    field = field.copy()
    pol = src.polarization
    if not is_arr:
        field[(pol, *gs)] += sign * (scale * F(inv_mat))
        return field
    sl = inv_mat[(slice(None), *gs)]
    if sl.shape[0] == 9:
        for axis in range(3):
            field[(axis, *gs)] += sign * (scale * sl[axis * 3 + pol])
        return field
    comp = sl[0] if sl.shape[0] == 1 else sl[pol]
    field[(pol, *gs)] += sign * (scale * comp)
    return field


def _apply_sources(field, arrays, objects, config, time_step: int, which: str, inverse: bool):
    """Source loop of update.py:494-519 / 824-849 (forward) and :559-584 / 878-903 (reverse)."""
    half = F(0.5) if which == "H" else F(0.0)
    for src in objects.sources:
        if src.uses_default_switch:
            t_f = F(time_step) + half
        else:
            if not bool(src._is_on_at_time_step_arr[time_step]):
                continue
            t_f = src.adjusted_time_step(time_step) + half
        if isinstance(src, TFSFPlaneSource):
            if which == "E":
                field = tfsf_update_E(src, field, arrays.inv_permittivities, t_f, inverse, config)
            else:
                field = tfsf_update_H(src, field, arrays.inv_permeabilities, t_f, inverse, config)
        elif isinstance(src, PointDipoleSource):
            mat = arrays.inv_permittivities if which == "E" else arrays.inv_permeabilities
            field = dipole_update(src, field, mat, t_f, inverse, config, which)
        else:
            raise NotImplementedError(type(src))
    return field


def apply_boundary_post_E_update(E, objects):
    for b in objects.boundary_objects:
        if isinstance(b, PerfectElectricConductor):
            c1, c2 = b.tangential_components
            E[(c1, *b.grid_slice)] = 0
            E[(c2, *b.grid_slice)] = 0
    return E


def apply_boundary_post_H_update(H, objects):
    for b in objects.boundary_objects:
        if isinstance(b, PerfectMagneticConductor):
            c1, c2 = b.tangential_components
            H[(c1, *b.grid_slice)] = 0
            H[(c2, *b.grid_slice)] = 0
    return H


# ----------------------------------------------------------------------------------------------
# field updates  (fdtd/update.py:256-523, 526-686, 689-853, 856-1007)
# ----------------------------------------------------------------------------------------------
def update_E(time_step: int, arrays: ArrayContainer, objects, config, simulate_boundaries: bool) -> ArrayContainer:
    inv_eps = arrays.inv_permittivities
    sigma_E = arrays.electric_conductivity
    c = F(config.courant_number)
    H_pad = pad_fields_for_boundaries(arrays.fields.H, objects, config)
    curl, psi_E = curl_H(config, H_pad, arrays.fields.psi_E, objects, simulate_boundaries)
    arrays = arrays.aset("fields->psi_E", psi_E)
    E_old = arrays.fields.E
    eps_full = inv_eps.shape[0] == 9
    sig_full = sigma_E is not None and sigma_E.shape[0] == 9

    if not eps_full and not sig_full:
        factor = F(1)
        if sigma_E is not None:
            factor = F(1) - c * sigma_E * F(eta0) * inv_eps / F(2)
        E = factor * E_old + c * curl * inv_eps
        if arrays.fields.dispersive_P_curr is not None:
            P_curr, P_prev = arrays.fields.dispersive_P_curr, arrays.fields.dispersive_P_prev
            c1, c2, c3, c4 = arrays.dispersive_c1, arrays.dispersive_c2, arrays.dispersive_c3, arrays.dispersive_c4
            P_hat = c1 * P_curr + c2 * P_prev + c3 * E_old
            delta_hat = _sum_poles(P_curr - P_hat)
            E = E + inv_eps * delta_hat
            if c4 is not None:
                divisor = F(1) + inv_eps * _sum_poles(c4)
                if sigma_E is not None:
                    divisor = divisor + c * sigma_E * F(eta0) * inv_eps / F(2)
                E = E / divisor
                P_new = P_hat + c4 * E
                arrays = arrays.aset("fields->dispersive_P_prev", P_curr)
                arrays = arrays.aset("fields->dispersive_P_curr", P_new.astype(F))
            else:
                arrays = arrays.aset("fields->dispersive_P_prev", P_curr)
                arrays = arrays.aset("fields->dispersive_P_curr", P_hat.astype(F))
                if sigma_E is not None:
                    E = E / (F(1) + c * sigma_E * F(eta0) * inv_eps / F(2))
        elif sigma_E is not None:
            E = E / (F(1) + c * sigma_E * F(eta0) * inv_eps / F(2))
    else:
        A, B = _update_matrices_cached(inv_eps, sigma_E, config.courant_number, eta0)
        if arrays.fields.dispersive_P_curr is not None:
            P_curr, P_prev = arrays.fields.dispersive_P_curr, arrays.fields.dispersive_P_prev
            c1, c2, c3 = arrays.dispersive_c1, arrays.dispersive_c2, arrays.dispersive_c3
            if arrays.dispersive_c4 is not None:
                raise AssertionError("CCPR poles are rejected for the full-tensor branch (update.py:406)")
            if c3.shape[1] == 9:
                raise NotImplementedError("oriented (9-component) dispersive coupling is out of scope")
            P_hat = c1 * P_curr + c2 * P_prev + c3 * E_old
            delta = _sum_poles(P_curr - P_hat)
            curl = curl + delta / c
            arrays = arrays.aset("fields->dispersive_P_prev", P_curr)
            arrays = arrays.aset("fields->dispersive_P_curr", P_hat.astype(F))
        E = _tensor_apply(E_old, curl, A, B, objects, config, "E", +1.0)

    E = _C(np.array(E))
    E = _apply_sources(E, arrays, objects, config, time_step, "E", inverse=False)
    E = apply_boundary_post_E_update(E, objects)
    return arrays.aset("fields->E", E)


def _sum_poles(x):
    """jnp.sum(axis=0): sequential float32 accumulation over the (small) pole axis."""
    acc = x[0].astype(F)
    for p in range(1, x.shape[0]):
        acc = acc + x[p]
    return acc


def update_E_reverse(time_step: int, arrays: ArrayContainer, objects, config) -> ArrayContainer:
    if arrays.fields.dispersive_P_curr is not None:
        raise NotImplementedError(
            "Dispersive time-reversible gradient computation under active development. "
            "Use GradientConfig(method='checkpointed') instead."
        )
    E = _apply_sources(arrays.fields.E, arrays, objects, config, time_step, "E", inverse=True)
    inv_eps = arrays.inv_permittivities
    sigma_E = arrays.electric_conductivity
    c = F(config.courant_number)
    H_pad = pad_fields_for_boundaries(arrays.fields.H, objects, config)
    curl, _ = curl_H(config, H_pad, arrays.fields.psi_E, objects, False)
    eps_full = inv_eps.shape[0] == 9
    sig_full = sigma_E is not None and sigma_E.shape[0] == 9
    if not eps_full and not sig_full:
        factor = F(1)
        if sigma_E is not None:
            E = E * (F(1) + c * sigma_E * F(eta0) * inv_eps / F(2))
            factor = F(1) - c * sigma_E * F(eta0) * inv_eps / F(2)
        E = (E - c * curl * inv_eps) / factor
    else:
        A, B = _update_matrices_cached(inv_eps, sigma_E, config.courant_number, eta0, reverse=True)
        E = _tensor_apply(E, curl, A, B, objects, config, "E", -1.0)
    E = apply_boundary_post_E_update(_C(np.array(E)), objects)
    return arrays.aset("fields->E", E)


def update_H(time_step: int, arrays: ArrayContainer, objects, config, simulate_boundaries: bool) -> ArrayContainer:
    inv_mu = arrays.inv_permeabilities
    sigma_H = arrays.magnetic_conductivity
    c = F(config.courant_number)
    E_pad = pad_fields_for_boundaries(arrays.fields.E, objects, config)
    curl, psi_H = curl_E(config, E_pad, arrays.fields.psi_H, objects, simulate_boundaries)
    arrays = arrays.aset("fields->psi_H", psi_H)
    mu_arr = isinstance(inv_mu, np.ndarray) and inv_mu.ndim > 0
    mu_full = mu_arr and inv_mu.shape[0] == 9
    sig_full = sigma_H is not None and sigma_H.shape[0] == 9
    mu = inv_mu if mu_arr else F(inv_mu)
    if not mu_full and not sig_full:
        factor = F(1)
        if sigma_H is not None:
            factor = F(1) - c * sigma_H / F(eta0) * mu / F(2)
        H = factor * arrays.fields.H - c * curl * mu
        if sigma_H is not None:
            H = H / (F(1) + c * sigma_H / F(eta0) * mu / F(2))
    else:
        A, B = _update_matrices_cached(inv_mu, sigma_H, config.courant_number, 1 / eta0)
        H = _tensor_apply(arrays.fields.H, curl, A, B, objects, config, "H", -1.0)
    H = _C(np.array(H))
    H = _apply_sources(H, arrays, objects, config, time_step, "H", inverse=False)
    H = apply_boundary_post_H_update(H, objects)
    return arrays.aset("fields->H", H)


def update_H_reverse(time_step: int, arrays: ArrayContainer, objects, config) -> ArrayContainer:
    H = _apply_sources(arrays.fields.H, arrays, objects, config, time_step, "H", inverse=True)
    inv_mu = arrays.inv_permeabilities
    sigma_H = arrays.magnetic_conductivity
    c = F(config.courant_number)
    E_pad = pad_fields_for_boundaries(arrays.fields.E, objects, config)
    curl, _ = curl_E(config, E_pad, arrays.fields.psi_H, objects, False)
    mu_arr = isinstance(inv_mu, np.ndarray) and inv_mu.ndim > 0
    mu_full = mu_arr and inv_mu.shape[0] == 9
    sig_full = sigma_H is not None and sigma_H.shape[0] == 9
    mu = inv_mu if mu_arr else F(inv_mu)
    if not mu_full and not sig_full:
        factor = F(1)
        if sigma_H is not None:
            H = H * (F(1) + c * sigma_H / F(eta0) * mu / F(2))
            factor = F(1) - c * sigma_H / F(eta0) * mu / F(2)
        H = (H + c * curl * mu) / factor
    else:
        A, B = _update_matrices_cached(inv_mu, sigma_H, config.courant_number, 1 / eta0, reverse=True)
        H = _tensor_apply(H, curl, A, B, objects, config, "H", +1.0)
    H = apply_boundary_post_H_update(_C(np.array(H)), objects)
    return arrays.aset("fields->H", H)


# ----------------------------------------------------------------------------------------------
# detector co-location stencil  (core/physics/curl.py:42-224)
# ----------------------------------------------------------------------------------------------
def _backward_edge_average(current, previous, config, axis: int, region_slice=None):
    if config is None or not config.has_nonuniform_grid:
        return F(0.5) * (current + previous)
    grid = config.resolved_grid
    widths = grid.cell_widths(axis)
    prev_widths = np.concatenate([widths[:1], widths[:-1]])
    if region_slice is not None:
        a, b = region_slice[axis]
        widths, prev_widths = widths[a:b], prev_widths[a:b]
    cur_hw = F(0.5) * widths
    prev_hw = F(0.5) * prev_widths
    bs = [1, 1, 1]
    bs[axis] = current.shape[axis]
    cur_hw, prev_hw = cur_hw.reshape(bs), prev_hw.reshape(bs)
    return (current * prev_hw + previous * cur_hw) / (cur_hw + prev_hw)


def interpolate_fields(E_pad, H_pad, config=None, region_slice=None):
    Ex, Ey, Ez = E_pad[0], E_pad[1], E_pad[2]
    Hx, Hy, Hz = H_pad[0], H_pad[1], H_pad[2]
    bea = lambda cur, prev, axis: _backward_edge_average(cur, prev, config, axis, region_slice)
    Ex_lo = bea(Ex[1:-1, 1:-1, 1:-1], Ex[:-2, 1:-1, 1:-1], 0)
    Ex_hi = bea(Ex[1:-1, 1:-1, 2:], Ex[:-2, 1:-1, 2:], 0)
    Ex_i = (Ex_lo + Ex_hi) / F(2.0)
    Ey_lo = bea(Ey[1:-1, 1:-1, 1:-1], Ey[1:-1, :-2, 1:-1], 1)
    Ey_hi = bea(Ey[1:-1, 1:-1, 2:], Ey[1:-1, :-2, 2:], 1)
    Ey_i = (Ey_lo + Ey_hi) / F(2.0)
    Ez_i = Ez[1:-1, 1:-1, 1:-1]
    Hx_i = bea(Hx[1:-1, 1:-1, 1:-1], Hx[1:-1, :-2, 1:-1], 1)
    Hy_i = bea(Hy[1:-1, 1:-1, 1:-1], Hy[:-2, 1:-1, 1:-1], 0)
    lo_x = bea(Hz[1:-1, 1:-1, 1:-1], Hz[:-2, 1:-1, 1:-1], 0)
    lo_xy = bea(lo_x, bea(Hz[1:-1, :-2, 1:-1], Hz[:-2, :-2, 1:-1], 0), 1)
    hi_x = bea(Hz[1:-1, 1:-1, 2:], Hz[:-2, 1:-1, 2:], 0)
    hi_xy = bea(hi_x, bea(Hz[1:-1, :-2, 2:], Hz[:-2, :-2, 2:], 0), 1)
    Hz_i = (lo_xy + hi_xy) / F(2.0)
    return _C(np.stack([Ex_i, Ey_i, Ez_i])), _C(np.stack([Hx_i, Hy_i, Hz_i]))


# ----------------------------------------------------------------------------------------------
# detectors  (objects/detectors/*.py, core/physics/metrics.py:15-117, fdtd/update.py:1040-1137)
# ----------------------------------------------------------------------------------------------
def compute_energy(E, H, inv_permittivity, inv_permeability):
    eps_shape = getattr(inv_permittivity, "shape", ())
    mu_shape = getattr(inv_permeability, "shape", ())
    if (eps_shape and eps_shape[0] == 9) or (mu_shape and mu_shape[0] == 9):
        perm, inv_perm = (2, 3, 4, 0, 1), (3, 4, 0, 1, 2)
        ie = expand_to_3x3(inv_permittivity)
        im = expand_to_3x3(inv_permeability)
        sp = E.shape[1:]
        ie = np.broadcast_to(ie, (3, 3, *sp))
        im = np.broadcast_to(im, (3, 3, *sp))
        eps = np.linalg.inv(ie.transpose(perm)).transpose(inv_perm).astype(F)
        mu = np.linalg.inv(im.transpose(perm)).transpose(inv_perm).astype(F)
        eE = F(0.5) * np.einsum("ixyz,ijxyz,jxyz->xyz", E, eps, E).astype(F)
        eH = F(0.5) * np.einsum("ixyz,ijxyz,jxyz->xyz", H, mu, H).astype(F)
        return (eE + eH).astype(F)
    E2 = np.square(np.abs(E))
    eE = F(0.5) * (F(1) / inv_permittivity) * E2
    eE = _sum0(np.broadcast_to(eE, E.shape))
    H2 = np.square(np.abs(H))
    eH = F(0.5) * (F(1) / (inv_permeability if isinstance(inv_permeability, np.ndarray) else F(inv_permeability))) * H2
    eH = _sum0(np.broadcast_to(eH, H.shape))
    return (eE + eH).astype(F)


def _sum0(x):
    return (x[0] + x[1]) + x[2]


def compute_poynting_flux(E, H):
    if np.iscomplexobj(E) or np.iscomplexobj(H):  # jnp.cross(E, conj(H)) (metrics.py:99-117); callers take .real
        H = np.conj(H)
        return np.stack([E[1] * H[2] - E[2] * H[1], E[2] * H[0] - E[0] * H[2], E[0] * H[1] - E[1] * H[0]], axis=0).real.astype(F)
    return np.stack(
        [E[1] * H[2] - E[2] * H[1], E[2] * H[0] - E[0] * H[2], E[0] * H[1] - E[1] * H[0]], axis=0
    ).astype(F)


def _select_components(det, E, H):
    fields = []
    for n, arr in zip(COMPONENT_NAMES, (E[0], E[1], E[2], H[0], H[1], H[2])):
        if n in det.components:
            fields.append(arr)
    return np.stack(fields, axis=0)


def detector_update(det, time_step: int, E, H, state: dict, inv_eps, inv_mu) -> dict:
    state = {k: v.copy() for k, v in state.items()}
    if isinstance(det, ClosedSurfacePhasorPoyntingFluxDetector):
        # poynting_flux.py:476-503: the two boundary planes of every active axis, no window factor
        time_passed = F(time_step) * F(det._dt)
        EH = np.stack([E[0], E[1], E[2], H[0], H[1], H[2]], axis=0)
        ang = (det._angular_frequencies * time_passed).astype(F)
        ph = (np.cos(ang) + 1j * np.sin(ang)).astype(np.complex64).reshape((len(ang),) + (1,) * EH.ndim)
        new = ((EH[None].astype(np.complex64) * ph).astype(np.complex64) * np.complex64(det._static_scale())).astype(np.complex64)
        for a in det._resolve_active_axes():
            for side, sl in (("min", slice(0, 1)), ("max", slice(-1, None))):
                idx = [slice(None)] * new.ndim
                idx[a + 2] = sl
                key = f"phasor_axis{a}_{side}"
                face = new[tuple(idx)][None]
                state[key] = (state[key] - face if det.inverse else state[key] + face).astype(np.complex64)
        return state
    if isinstance(det, ClosedSurfacePoyntingFluxDetector):
        # poynting_flux.py:263-284 + metrics.py:120-160
        pf = compute_poynting_flux(E, H)
        net = F(0.0)
        for a in det._resolve_active_axes():
            weighted = (pf[a] * det._face_area_weights_per_axis[a]).astype(F)
            net = net + np.take(weighted, -1, axis=a).sum(dtype=F) - np.take(weighted, 0, axis=a).sum(dtype=F)
        if det.orientation == "inward":
            net = -net
        state["poynting_flux"][int(det._time_step_to_arr_idx[time_step])] = F(net)
        return state
    if isinstance(det, PhasorDetector):
        time_passed = F(time_step) * F(det._dt)
        scale = det._static_scale()
        w = det._window_at_time_step_arr[time_step]
        EH = _select_components(det, E, H)
        ang = (det._angular_frequencies * time_passed).astype(F)
        ph = (np.cos(ang) + 1j * np.sin(ang)).astype(np.complex64)
        ph = ph.reshape((len(ang),) + (1,) * EH.ndim)
        new = (EH[None].astype(np.complex64) * ph).astype(np.complex64)
        new = (new * np.complex64(scale)).astype(np.complex64)
        new = (new * np.complex64(w)).astype(np.complex64)
        if det.reduce_volume:
            wts = det._cached_cell_volume_weights
            new = ((new * wts[None, None]).sum(axis=(2, 3, 4), dtype=np.complex64) / wts.sum(dtype=F)).astype(np.complex64)
        if det.inverse:
            state["phasor"] = (state["phasor"] - new[None]).astype(np.complex64)
        else:
            state["phasor"] = (state["phasor"] + new[None]).astype(np.complex64)
        return state
    idx = int(det._time_step_to_arr_idx[time_step])
    if isinstance(det, EnergyDetector):
        energy = compute_energy(E, H, inv_eps, inv_mu)
        if det.as_slices:
            if det.use_mean:
                state["XY Plane"][idx] = energy.mean(axis=2, dtype=F)
                state["XZ Plane"][idx] = energy.mean(axis=1, dtype=F)
                state["YZ Plane"][idx] = energy.mean(axis=0, dtype=F)
            else:
                xi, yi, zi = det._slice_indices
                state["XY Plane"][idx] = energy[:, :, zi]
                state["XZ Plane"][idx] = energy[:, yi, :]
                state["YZ Plane"][idx] = energy[xi, :, :]
        elif det.reduce_volume:
            state["energy"][idx] = (energy * det._cached_cell_volume_weights).sum(dtype=F)
        else:
            state["energy"][idx] = energy
        return state
    if isinstance(det, FieldDetector):
        EH = _select_components(det, E, H)
        if det.reduce_volume:
            wts = det._cached_cell_volume_weights
            EH = (EH * wts[None]).sum(axis=(1, 2, 3), dtype=F) / wts.sum(dtype=F)
        state["fields"][idx] = EH
        return state
    if isinstance(det, PoyntingFluxDetector):
        pf = compute_poynting_flux(E, H)
        if not det.keep_all_components:
            pf = pf[det.propagation_axis]
        if det.direction == "-":
            pf = -pf
        if det.reduce_volume:
            pf = pf * det._cached_face_area_weights
            pf = pf.sum(axis=(1, 2, 3), dtype=F) if det.keep_all_components else pf.sum(dtype=F)
        state["poynting_flux"][idx] = pf
        return state
    raise NotImplementedError(type(det))


def update_detector_states(time_step: int, arrays: ArrayContainer, objects, config, H_prev, inverse: bool):
    to_update = objects.backward_detectors if inverse else objects.forward_detectors
    if not to_update:
        return arrays
    state = dict(arrays.detector_states)
    grid_shape = objects.volume.grid_shape
    E, H = arrays.fields.E, arrays.fields.H

    def is_interior(d):
        return all(s >= 1 and e <= grid_shape[a] - 1 for a, (s, e) in enumerate(d.grid_slice_tuple))

    full = None
    for d in to_update:
        if not bool(d._is_on_at_time_step_arr[time_step]):
            continue
        gs = d.grid_slice
        if not d.exact_interpolation:
            E_reg, H_reg = E[(slice(None), *gs)], H[(slice(None), *gs)]
        elif is_interior(d):
            block = (slice(None), *(slice(s - 1, e + 1) for (s, e) in d.grid_slice_tuple))
            H_avg = (H_prev[block] + H[block]) / F(2)
            E_reg, H_reg = interpolate_fields(E[block], H_avg, config=config, region_slice=d.grid_slice_tuple)
        else:
            if full is None:
                full = interpolate_fields(
                    pad_fields_with_symmetry_mirror(E, objects, config, "E"),
                    pad_fields_with_symmetry_mirror((H_prev + H) / F(2), objects, config, "H"),
                    config=config,
                )
            E_reg, H_reg = full[0][(slice(None), *gs)], full[1][(slice(None), *gs)]
        inv_mu = arrays.inv_permeabilities
        state[d.name] = detector_update(
            d,
            time_step,
            E_reg,
            H_reg,
            state[d.name],
            arrays.inv_permittivities[(slice(None), *gs)],
            inv_mu[(slice(None), *gs)] if isinstance(inv_mu, np.ndarray) and inv_mu.ndim > 0 else inv_mu,
        )
    return arrays.aset("detector_states", state)


# ----------------------------------------------------------------------------------------------
# PML-interface record / replay  (fdtd/misc.py:10-66, fdtd/update.py:1140-1222, interfaces/*)
# ----------------------------------------------------------------------------------------------
def _cast_to_rec(x: np.ndarray, recorder):
    if recorder.dtype_code == REC_F32:
        return _C(x)
    import torch

    return _TorchLeaf(torch.from_numpy(np.ascontiguousarray(x)).to(recorder.torch_dtype()))


def _cast_from_rec(x) -> np.ndarray:
    if isinstance(x, _TorchLeaf):
        import torch

        return x.t.to(torch.float32).numpy()
    return _C(x)


def _rec_get(buf, idx: int) -> np.ndarray:
    if isinstance(buf, _TorchLeaf):
        return _cast_from_rec(_TorchLeaf(buf.t[idx]))
    return _C(buf[idx])


def _rec_set(buf, idx: int, val: np.ndarray, recorder):
    v = _cast_to_rec(val, recorder)
    if isinstance(buf, _TorchLeaf):
        buf.t[idx] = v.t
    else:
        buf[idx] = v


def collect_interfaces(time_step: int, arrays: ArrayContainer, objects, config) -> ArrayContainer:
    if config.gradient_config is None or config.gradient_config.recorder is None:
        raise Exception("Need recorder to record boundaries")
    if arrays.recording_state is None:
        raise Exception("Need recording state to record boundaries")
    rec = config.gradient_config.recorder
    slot = int(rec.slot_of_time[time_step])
    if slot < 0:
        return arrays
    data = arrays.recording_state.data
    for fs in ("E", "H"):
        arr = getattr(arrays.fields, fs)
        for pml in objects.pml_objects:
            _rec_set(data[f"{pml.name}_{fs}"], slot, arr[(slice(None), *pml.interface_slice())], rec)
    return arrays


def add_interfaces(time_step: int, arrays: ArrayContainer, objects, config) -> ArrayContainer:
    if config.gradient_config is None or config.gradient_config.recorder is None:
        raise Exception("Need recorder to record boundaries")
    if arrays.recording_state is None:
        raise Exception("Need recording state to record boundaries")
    rec = config.gradient_config.recorder
    a, b, w = int(rec.replay_a[time_step]), int(rec.replay_b[time_step]), F(rec.replay_w[time_step])
    data = arrays.recording_state.data
    E, H = arrays.fields.E.copy(), arrays.fields.H.copy()
    for fs, arr in (("E", E), ("H", H)):
        for pml in objects.pml_objects:
            buf = data[f"{pml.name}_{fs}"]
            prev = _rec_get(buf, a)
            if a == b:
                val = prev
            else:
                nxt = _rec_get(buf, b)
                val = prev + w * (nxt - prev)
            arr[(slice(None), *pml.interface_slice())] = val
    arrays = arrays.aset("fields->E", E)
    return arrays.aset("fields->H", H)


# ----------------------------------------------------------------------------------------------
# one step forward / backward, loop drivers  (fdtd/forward.py:83-156, backward.py:18-135,
# fdtd/fdtd.py:421-584)
# ----------------------------------------------------------------------------------------------
def forward(state, config, objects, key=None, record_detectors=True, record_boundaries=False, simulate_boundaries=True):
    time_step, arrays = state
    time_step = int(time_step)
    H_prev = arrays.fields.H
    arrays = update_E(time_step, arrays, objects, config, simulate_boundaries)
    arrays = update_H(time_step, arrays, objects, config, simulate_boundaries)
    if record_boundaries:
        arrays = collect_interfaces(time_step, arrays, objects, config)
    if record_detectors:
        arrays = update_detector_states(time_step, arrays, objects, config, H_prev, inverse=False)
    return (time_step + 1, arrays)


def backward(state, config, objects, key=None, record_detectors=True, reset_fields=True, fields_to_reset=("E", "H")):
    time_step, arrays = state
    time_step = int(time_step) - 1
    arrays = add_interfaces(time_step, arrays, objects, config)
    H = arrays.fields.H
    arrays = update_H_reverse(time_step, arrays, objects, config)
    arrays = update_E_reverse(time_step, arrays, objects, config)
    if reset_fields:
        for name in fields_to_reset:
            f = getattr(arrays.fields, name).copy()
            for b in objects.boundary_objects:
                if isinstance(b, PerfectlyMatchedLayer):
                    f[(slice(None), *b.grid_slice)] = 0
                # BlochBoundary.apply_field_reset copies its own face onto itself: a no-op
            arrays = arrays.aset(f"fields->{name}", f)
    if record_detectors:
        arrays = update_detector_states(time_step, arrays, objects, config, H, inverse=True)
    return (time_step, arrays)


def custom_fdtd_forward(arrays, objects, config, key=None, reset_container=True, record_detectors=True, start_time=0, end_time=0, record_boundaries=False):
    if reset_container:
        arrays = arrays.reset()
    state = (int(start_time), arrays)
    while end_time > state[0]:
        state = forward(state, config, objects, key, record_detectors, record_boundaries, True)
    return state


def evaluate_condition(cond, state, config, objects) -> bool:
    """CPU restatement of the reference's stopping conditions (fdtd/stop_conditions.py:65-78,
    124-147, 269-332) on NumPy containers; ``cond`` is a set-up ``fdtdx_b200.stop_conditions`` object
    (attributes only - none of its device code runs here)."""
    name = type(cond).__name__
    t, arrays = state
    if name == "TimeStepCondition":
        return bool(config.time_steps_total > t)
    if name == "EnergyThresholdCondition":
        time_condition = t < cond.max_steps
        min_steps_condition = t < cond.min_steps
        total_energy = np.sum(compute_energy(arrays.fields.E, arrays.fields.H, arrays.inv_permittivities, arrays.inv_permeabilities))
        converged = total_energy < cond.threshold
        return bool(time_condition & (min_steps_condition | (not converged)))
    if name == "DetectorConvergenceCondition":
        spp, pp, total = cond._spp, cond.prev_periods, config.time_steps_total
        readings = next(iter(arrays.detector_states[cond.detector_name].values()))
        time_condition = t < total
        min_steps_condition = t >= cond.min_steps
        converged = False
        if min_steps_condition:
            start_ref = int(np.clip(t - (pp + 1) * spp, 0, total - pp * spp))
            start_last = int(np.clip(t - spp, 0, total - spp))
            ref = readings[start_ref : start_ref + pp * spp, 0].astype(F)
            last = readings[start_last : start_last + spp, 0].astype(F)
            ref_mean = np.mean(ref.reshape(pp, spp), axis=0, dtype=F)
            distance = np.linalg.norm(np.abs(np.fft.rfft(ref_mean, n=spp)) - np.abs(np.fft.rfft(last, n=spp)))
            converged = bool(distance < cond.threshold)
        return bool((not min_steps_condition) | (time_condition & (not converged)))
    raise NotImplementedError(name)


def checkpointed_fdtd(arrays, objects, config, key=None, stopping_condition=None):
    """fdtd.py:421-496: ``while cond(state): state = forward(state)`` bounded by time_steps_total."""
    arrays = arrays.reset()
    state = (0, arrays)
    cond = None if stopping_condition is None else stopping_condition.setup(state, config, objects)
    while state[0] < config.time_steps_total and (cond is None or evaluate_condition(cond, state, config, objects)):
        state = forward(state, config, objects, key, True, config.invertible_optimization, True)
    return state


def run_fdtd(arrays, objects, config, key=None, stopping_condition=None):
    return checkpointed_fdtd(arrays, objects, config, key, stopping_condition)


def full_backward(state, objects, config, key=None, record_detectors=True, reset_fields=True, start_time_step=0):
    while state[0] > start_time_step:
        state = backward(state, config, objects, key, record_detectors, reset_fields)
    return state
