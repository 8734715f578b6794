"""Waveguide mode solver and mode-source / mode-overlap setup.

The reference delegates the eigen-solve to a third-party dependency that is absent here:
``tidy3d>=2.8.0`` (``pyproject.toml:19``), ``tidy3d.components.mode.solver.compute_modes`` called from
``fdtdx/core/physics/modes.py:579-728``.  Its published algorithm (the diagonal-tensor branch,
``solver_diagonal``) is restated below with SciPy sparse matrices:

* the cross-section is a 2-D Yee grid, coordinates scaled by ``k0``; forward differences act on E,
  backward differences on H; walls are PEC (default) or PMC at the min edge of an axis;
* with ``P = [[-Dxf e_zz^-1 Dyb, Dxf e_zz^-1 Dxb + mu_yy], [-Dyf e_zz^-1 Dyb - mu_xx, Dyf e_zz^-1 Dxb]]`` and
  ``Q = [[-Dxb m_zz^-1 Dyf, Dxb m_zz^-1 Dxf + e_yy], [-Dyb m_zz^-1 Dyf - e_xx, Dyb m_zz^-1 Dxf]]`` the transverse
  E field solves ``(P Q) [Ex; Ey] = -(n_eff + i k_eff)^2 [Ex; Ey]`` (shift-invert Arnoldi around the
  largest permittivity);
* ``[Hx; Hy] = Q [Ex; Ey] / (i n_eff - k_eff)``, ``Hz = m_zz^-1 (Dxf Ey - Dyf Ex)``,
  ``Ez = e_zz^-1 (Dxb Hy - Dyb Hx)``, then ``H *= -i / eta0``.

Around it, ``compute_mode`` mirrors ``core/physics/modes.py:101-385`` (axis permutation to the solver's
"propagation along z" frame, polarisation filter ``:38-98``, H back to field units, Poynting
normalisation ``metrics.py:163-219``), ``make_mode_source`` mirrors ``objects/sources/mode.py:95-276``
and ``mode_overlap`` the overlap integral of ``objects/detectors/mode.py:379-423``.
Host-side setup (NumPy / SciPy, complex128); nothing here runs in the time loop.
"""

from __future__ import annotations

import numpy as np

from fdtdx_b200 import constants
from fdtdx_b200.sources import TFSFPlaneSource, calculate_time_offset_yee
from fdtdx_b200.switch import OnOffSwitch

_f32 = np.float32


def _diff_matrices(shape, dl_f, dl_b, pmc):
    """Forward (E) / backward (H) difference operators of the 2-D cross-section, C-order flattening
    (index = ix * Ny + iy).  PEC wall: the first forward row keeps only the +1 entry and the first
    backward row vanishes; PMC wall: the first backward row doubles."""
    import scipy.sparse as sp

    nx, ny = shape

    def one(n, widths, forward, is_pmc):
        if n == 1:
            return sp.csr_matrix((1, 1))
        if forward:
            d = sp.lil_matrix(sp.diags([-1.0, 1.0], [0, 1], shape=(n, n)))
            if not is_pmc:
                d[0, 0] = 0.0
        else:
            d = sp.lil_matrix(sp.diags([1.0, -1.0], [0, -1], shape=(n, n)))
            d[0, 0] = 2.0 if is_pmc else 0.0
        return sp.diags(1.0 / widths) @ sp.csr_matrix(d)

    ix, iy = sp.eye(nx), sp.eye(ny)
    dxf = sp.kron(one(nx, dl_f[0], True, pmc[0]), iy, format="csr")
    dxb = sp.kron(one(nx, dl_b[0], False, pmc[0]), iy, format="csr")
    dyf = sp.kron(ix, one(ny, dl_f[1], True, pmc[1]), format="csr")
    dyb = sp.kron(ix, one(ny, dl_b[1], False, pmc[1]), format="csr")
    return dxf, dxb, dyf, dyb


def solve_modes_diagonal(eps, mu, coords, frequency: float, num_modes: int, symmetry=(0, 0), target_neff: float | None = None):
    """Modes of a cross-section with diagonal material tensors.

    ``eps`` / ``mu``: (3, n0, n1) relative tensors' diagonals in the solver frame (x, y transverse, z
    propagation); ``coords``: two edge-coordinate arrays in metres.  Returns ``(E, H, neff)`` with E, H
    of shape (3, n0, n1, num_modes) (H in SI-normalised units, i.e. divided by eta0) and complex
    ``neff`` sorted by decreasing real part."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    eps = np.asarray(eps, np.complex128)
    mu = np.asarray(mu, np.complex128)
    n0, n1 = eps.shape[1:]
    N = n0 * n1
    k0 = 2.0 * np.pi * frequency / constants.c
    sc = [np.asarray(c, np.float64) * k0 for c in coords]
    dl_f = [c[1:] - c[:-1] for c in sc]
    dl_b = [np.concatenate([d[:1], 0.5 * (d[:-1] + d[1:])]) for d in dl_f]
    dxf, dxb, dyf, dyb = _diff_matrices((n0, n1), dl_f, dl_b, (symmetry[0] == 1, symmetry[1] == 1))
    dg = lambda v: sp.diags(v.reshape(-1))
    e_xx, e_yy, e_zz = eps
    m_xx, m_yy, m_zz = mu
    ie_zz, im_zz = dg(1.0 / e_zz), dg(1.0 / m_zz)
    P = sp.bmat([[-dxf @ ie_zz @ dyb, dxf @ ie_zz @ dxb + dg(m_yy)], [-dyf @ ie_zz @ dyb - dg(m_xx), dyf @ ie_zz @ dxb]])
    Q = sp.bmat([[-dxb @ im_zz @ dyf, dxb @ im_zz @ dxf + dg(e_yy)], [-dyb @ im_zz @ dyf - dg(e_xx), dyb @ im_zz @ dxf]]).tocsr()
    A = (P @ Q).tocsc()
    lossless = bool(np.all(eps.imag == 0) and np.all(mu.imag == 0))
    if lossless:
        A = A.real
    guess = target_neff if target_neff is not None else float(np.sqrt(np.max(eps.real)))
    k = min(num_modes, 2 * N - 2)
    rng = np.random.default_rng(0)  # deterministic start vector -> reproducible eigenvector phases
    vals, vecs = spl.eigs(A, k=k, sigma=-(guess**2), v0=rng.random(2 * N), which="LM")
    nk = np.sqrt(-vals.astype(np.complex128))
    nk = np.where(nk.real < 0, -nk, nk)
    order = np.argsort(-nk.real, kind="stable")
    nk, vecs = nk[order], vecs[:, order].astype(np.complex128)
    Ex, Ey = vecs[:N], vecs[N:]
    h = Q @ vecs
    Hx, Hy = h[:N] / (1j * nk)[None, :], h[N:] / (1j * nk)[None, :]
    Hz = im_zz @ (dxf @ Ey - dyf @ Ex)
    Ez = ie_zz @ (dxb @ Hy - dyb @ Hx)
    E = np.stack([Ex, Ey, Ez]).reshape(3, n0, n1, k)
    H = np.stack([Hx, Hy, Hz]).reshape(3, n0, n1, k) * (-1j / constants.eta0)
    # fix the arbitrary eigenvector phase: the largest transverse E sample of each mode is real positive
    for m in range(k):
        t = E[:2, :, :, m].reshape(-1)
        ph = t[np.argmax(np.abs(t))]
        ph = ph / abs(ph) if abs(ph) > 0 else 1.0
        E[..., m] /= ph
        H[..., m] /= ph
    return E, H, nk


def _te_fraction(Ex, Ey, pol):
    a, b = np.sum(np.abs(Ex) ** 2), np.sum(np.abs(Ey) ** 2)
    return (a if pol == "te" else b) / (a + b + 1e-18)


def compute_mode(frequency: float, inv_permittivities, inv_permeabilities=1.0, resolution: float | None = None, direction: str = "+",
                 mode_index: int = 0, filter_pol: str | None = None, transverse_coords=None, symmetry=(0, 0)):
    """``core/physics/modes.py:101-385``.  ``inv_permittivities``: (1|3, nx, ny, nz) with exactly one
    singleton spatial axis (the propagation axis).  Returns ``(E, H, neff)``: complex (3, nx, ny, nz)
    fields in physical axis order, H in units of E (multiplied by eta0), Poynting-normalised."""
    inv_eps = np.asarray(inv_permittivities)
    if inv_eps.ndim != 4 or inv_eps.shape[0] not in (1, 3) or sum(d == 1 for d in inv_eps.shape[1:]) != 1:
        raise Exception(f"Invalid shape of inv_permittivities: {inv_eps.shape}")
    p = inv_eps.shape[1:].index(1)
    t0, t1 = [a for a in range(3) if a != p]
    eps = 1.0 / np.take(inv_eps.astype(np.complex128), 0, axis=p + 1)
    eps = np.repeat(eps, 3, axis=0) if eps.shape[0] == 1 else eps
    mu_in = inv_permeabilities
    if hasattr(mu_in, "shape") and np.ndim(mu_in) > 0:
        mu = 1.0 / np.take(np.asarray(mu_in, np.complex128), 0, axis=p + 1)
        mu = np.repeat(mu, 3, axis=0) if mu.shape[0] == 1 else mu
    else:
        mu = np.full(eps.shape, 1.0 / complex(mu_in))
    # solver frame (x, y, z) = physical (t0, t1, p)
    perm = [t0, t1, p]
    eps_s, mu_s = eps[perm], mu[perm]
    if transverse_coords is None:
        if resolution is None:
            raise ValueError("resolution is required when transverse_coords is not provided")
        coords = [np.arange(eps_s.shape[1] + 1) * resolution, np.arange(eps_s.shape[2] + 1) * resolution]
        area = None
    else:
        coords = [np.asarray(c, np.float64) for c in transverse_coords]
        area = np.diff(coords[0])[:, None] * np.diff(coords[1])[None, :]
    E, H, neff = solve_modes_diagonal(eps_s, mu_s, coords, frequency, 2 * (mode_index + 1) + 10, symmetry)
    idx = list(range(E.shape[-1]))
    if filter_pol is not None:
        match = [m for m in idx if _te_fraction(E[0, ..., m], E[1, ..., m], filter_pol) >= 0.5]
        idx = match + [m for m in idx if m not in match]
    m = idx[mode_index]
    Es, Hs, n = E[..., m], H[..., m], neff[m]
    if direction == "-":  # reciprocity: E_z -> -E_z, H_t -> -H_t  (modes.py:699-701)
        Es = np.stack([Es[0], Es[1], -Es[2]])
        Hs = np.stack([-Hs[0], -Hs[1], Hs[2]])
    # back to physical component order (modes.py:217-232); the p == 1 frame is left-handed -> H flips
    inv = [perm.index(a) for a in range(3)]
    Ep, Hp = Es[inv], Hs[inv]
    if p == 1:
        Hp = -Hp
    Hp = Hp * constants.eta0
    Ep, Hp = np.expand_dims(Ep, p + 1), np.expand_dims(Hp, p + 1)
    # normalize_by_poynting_flux (metrics.py:163-219)
    S = np.cross(np.conj(Ep), Hp, axisa=0, axisb=0, axisc=0)[p]
    Sr = 0.5 * S.real
    if area is not None:
        Sr = Sr * np.expand_dims(area / area.mean(), p)
    norm = np.sqrt(abs(Sr.sum()))
    return Ep / norm, Hp / norm, complex(n)


def _transverse_edges(config, slice_tuple, p):
    grid = config.resolved_grid
    if grid is None or not config.has_nonuniform_grid:
        return None
    out = []
    for a in range(3):
        if a == p:
            continue
        lo, hi = slice_tuple[a]
        e = grid.edges(a)[lo : hi + 1]
        out.append(np.asarray(e - e[0], np.float64))
    return out


def make_mode_source(name: str, grid_slice_tuple, config, inv_permittivities, inv_permeabilities=1.0, *, direction: str = "+",
                     wave_character, temporal_profile=None, mode_index: int = 0, filter_pol: str | None = None,
                     static_amplitude_factor: float = 1.0, switch: OnOffSwitch | None = None, electric_conductivity=None, inv_eps_slice=None):
    """``ModePlaneSource.apply`` (``objects/sources/mode.py:95-276``): solve the cross-section's mode,
    keep the real part of a lossless mode (the complex profile of a lossy one: quadrature injection,
    ``tfsf.py:266-283``), per-component Yee time offsets from ``Re(n_eff)``."""
    from fdtdx_b200.profile import SingleFrequencyProfile

    src = TFSFPlaneSource(name=name, grid_slice_tuple=grid_slice_tuple, wave_character=wave_character,
                          temporal_profile=temporal_profile or SingleFrequencyProfile(), static_amplitude_factor=static_amplitude_factor,
                          switch=switch or OnOffSwitch(), direction=direction)
    src.place_on_grid(config)
    p = src.propagation_axis
    gs = src.grid_slice
    # ``inv_eps_slice``: the (C, *face) cross-section itself, for callers that never materialise the volume
    inv_eps = np.asarray(inv_eps_slice) if inv_eps_slice is not None else np.asarray(inv_permittivities)[(slice(None), *gs)]
    mode_inv_eps = inv_eps
    sigma = None if electric_conductivity is None else np.asarray(electric_conductivity)[(slice(None), *gs)]
    if sigma is not None:
        # effective_complex_inv_permittivity (dispersion.py:1310-): eps + i sigma / (eps0 omega); the stored
        # conductivity is pre-multiplied by c0*dt/courant
        spacing = constants.c * config.time_step_duration / config.courant_number
        omega = 2.0 * np.pi * wave_character.get_frequency()
        eps_c = 1.0 / inv_eps.astype(np.complex128) + 1j * (sigma / spacing) / (constants.eps0 * omega)
        mode_inv_eps = 1.0 / eps_c
    mu = inv_permeabilities
    if hasattr(mu, "shape") and np.ndim(mu) > 0:
        mu = np.asarray(mu)[(slice(None), *gs)]
    spacing = None if config.has_nonuniform_grid else config.uniform_spacing()
    E, H, neff = compute_mode(wave_character.get_frequency(), mode_inv_eps, mu, resolution=spacing, direction=direction, mode_index=mode_index,
                              filter_pol=filter_pol, transverse_coords=_transverse_edges(config, grid_slice_tuple, p))
    if sigma is None:
        src._E, src._H = E.real.astype(_f32), H.real.astype(_f32)
    else:
        src._E, src._H = E.astype(np.complex64), H.astype(np.complex64)
    src._neff = neff
    face = src.grid_shape
    k = np.zeros(3, _f32)
    k[p] = 1.0 if direction == "+" else -1.0
    center = [0.0, 0.0, 0.0]  # k has no transverse component: the reference point does not matter
    src._time_offset_E, src._time_offset_H = calculate_time_offset_yee(center, k, np.full(face, _f32(neff.real), _f32), face, config, grid_slice_tuple)
    return src


def mode_overlap(phasor, mode_E, mode_H, axis: int, area_weights=None, pulse: bool = False):
    """``ModeOverlapDetector.compute_overlap_to_mode`` (``objects/detectors/mode.py:379-423``):
    ``alpha = sum( (E_m x conj(H) + conj(E) x H_m) . n * area )``, divided by 4 unless the detector uses
    pulse scaling.  ``phasor``: (6, *plane) complex (Ex..Hz of one frequency)."""
    E, H = phasor[:3], phasor[3:]
    a = np.cross(mode_E, np.conj(H), axisa=0, axisb=0, axisc=0)[axis] + np.cross(np.conj(E), mode_H, axisa=0, axisb=0, axisc=0)[axis]
    if area_weights is not None:
        a = a * area_weights
    alpha = a.sum()
    return complex(alpha if pulse else alpha / 4.0)
