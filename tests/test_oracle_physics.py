"""Physics oracles of the reference's simulation tier, re-run against the CPU oracle (SURVEY.md
section 8c): plane-wave phase velocity and impedance (tests/simulation/physics/test_plane_wave.py),
Fresnel T = 8/9, R = 1/9 (test_fresnel.py), PEC reflection R ~ 1 (boundaries/test_pec_reflection.py),
skin-depth decay (test_skin_depth.py), Lorentz-medium phase velocity (test_dispersion.py) and
1-step time reversal for every boundary kind (fdtd/test_time_reversal.py:251-327).  Tolerances are
the reference's (5 %; reversal atol 1e-5)."""

import math

import numpy as np
import pytest

import fdtdx_b200 as fx
from fdtdx_b200.constants import c as c0
from oracle import yee

F = np.float32
DX = 25e-9
WL = 1.0e-6


def _line_scene(nz, eps_fn=None, bnd_z=("pml", "pml"), time=None, detectors=(), sigma_fn=None, dispersive=None, src_z=14):
    shape = (3, 3, nz)
    cfg = fx.SimulationConfig(time=time or 30e-15, grid=fx.UniformGrid(spacing=DX))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    types = {"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": bnd_z[0], "max_z": bnd_z[1]}
    bl = fx.boundary_objects_from_config(shape, cfg, types, thickness=10)
    z = np.arange(nz)
    eps = np.ones(nz) if eps_fn is None else eps_fn(z)
    inv_eps = np.broadcast_to((1.0 / eps).astype(F)[None, None, None, :], (1, *shape)).copy()
    sig = None
    if sigma_fn is not None:
        ref = c0 * cfg.time_step_duration / cfg.courant_number
        sig = np.broadcast_to((sigma_fn(z) * ref).astype(F)[None, None, None, :], (1, *shape)).copy()
    src = fx.make_plane_source("s", ((0, 3), (0, 3), (src_z, src_z + 1)), cfg, inv_eps, direction="+", wave_character=fx.WaveCharacter(wavelength=WL), normalize_by_energy=False)
    objs = [vol, *bl, src, *[d.place_on_grid(cfg) for d in detectors]]
    objects, arrays, _, cfg, _ = fx.place_objects(objs, cfg, inv_permittivities=inv_eps, electric_conductivity=sig, dispersive=dispersive)
    return objects, arrays, cfg


def _point(name, z, comps=("Ex", "Hy")):
    return fx.FieldDetector(name=name, grid_slice_tuple=((1, 2), (1, 2), (z, z + 1)), components=comps, exact_interpolation=False)


def _fit_phase(sig, dt, omega):
    t = np.arange(sig.shape[0]) * dt
    a = 2 * np.mean(sig * np.exp(1j * omega * t))
    return a


def test_plane_wave_phase_velocity_and_impedance():
    n_med = 1.5
    nz = 140
    objects, arrays, cfg = _line_scene(nz, eps_fn=lambda z: np.full(z.shape, n_med**2), time=60e-15, detectors=[_point("a", 60), _point("b", 68)])
    T = cfg.time_steps_total
    st = yee.checkpointed_fdtd(arrays, objects, cfg)
    a = st[1].detector_states["a"]["fields"][:, :, 0, 0, 0]
    b = st[1].detector_states["b"]["fields"][:, :, 0, 0, 0]
    period_steps = WL / c0 / cfg.time_step_duration
    n_last = int(4 * period_steps)
    omega = 2 * np.pi * c0 / WL
    pa = _fit_phase(a[-n_last:, 0], cfg.time_step_duration, omega)
    pb = _fit_phase(b[-n_last:, 0], cfg.time_step_duration, omega)
    dphi = np.angle(pb / pa)
    k_meas = abs(dphi) / (8 * DX)
    k_expected = 2 * np.pi * n_med / WL
    assert abs(k_meas - k_expected) / k_expected < 0.05
    # impedance in the solver's normalised units: E/H = 1/n
    z_meas = np.abs(a[-n_last:, 0]).max() / np.abs(a[-n_last:, 1]).max()
    assert abs(z_meas - 1 / n_med) * n_med < 0.05


def test_fresnel_normal_incidence():
    """n = 1 -> 2: T = 8/9, R = 1/9 with two-run normalisation."""
    nz, zi = 160, 90
    mk = lambda name, z: fx.PoyntingFluxDetector(name=name, grid_slice_tuple=((0, 3), (0, 3), (z, z + 1)), direction="+", exact_interpolation=True)
    flux = {}
    for label, eps_fn in (("ref", None), ("slab", lambda z: np.where(z >= zi, 4.0, 1.0))):
        objects, arrays, cfg = _line_scene(nz, eps_fn=eps_fn, time=70e-15, detectors=[mk("front", 50), mk("back", 120)])
        st = yee.checkpointed_fdtd(arrays, objects, cfg)
        n_last = int(3 * WL / c0 / cfg.time_step_duration)
        flux[label] = {k: st[1].detector_states[k]["poynting_flux"][-n_last:, 0].mean() for k in ("front", "back")}
    T = flux["slab"]["back"] / flux["ref"]["back"]
    R = 1 - flux["slab"]["front"] / flux["ref"]["front"]
    assert abs(T - 8 / 9) < 0.05 and abs(R - 1 / 9) < 0.05


def test_pec_reflection_standing_wave():
    """A PEC wall reflects everything: net time-averaged flux in front of it is ~0."""
    nz = 120
    det = fx.PoyntingFluxDetector(name="p", grid_slice_tuple=((0, 3), (0, 3), (60, 61)), direction="+")
    ref_objects, ref_arrays, cfg = _line_scene(nz, time=60e-15, detectors=[det])
    ref = yee.checkpointed_fdtd(ref_arrays, ref_objects, cfg)[1].detector_states["p"]["poynting_flux"][:, 0]
    det2 = fx.PoyntingFluxDetector(name="p", grid_slice_tuple=((0, 3), (0, 3), (60, 61)), direction="+")
    objects, arrays, cfg = _line_scene(nz, bnd_z=("pml", "pec"), time=60e-15, detectors=[det2])
    out = yee.checkpointed_fdtd(arrays, objects, cfg)
    got = out[1].detector_states["p"]["poynting_flux"][:, 0]
    n_last = int(3 * WL / c0 / cfg.time_step_duration)
    assert abs(got[-n_last:].mean()) < 0.05 * abs(ref[-n_last:].mean())
    wall = objects.pec_objects[0]
    assert np.all(out[1].fields.E[(0, *wall.grid_slice)] == 0) and np.all(out[1].fields.E[(1, *wall.grid_slice)] == 0)


def test_skin_depth_decay():
    """Lossy half-space: field envelope decays as exp(-alpha z) with alpha from the complex k."""
    nz, zi = 200, 60
    sigma = 2.0e4  # S/m
    objects, arrays, cfg = _line_scene(nz, time=80e-15, sigma_fn=lambda z: np.where(z >= zi, sigma, 0.0),
                                       detectors=[_point(f"d{z}", z, ("Ex",)) for z in (80, 100, 120)])
    st = yee.checkpointed_fdtd(arrays, objects, cfg)
    n_last = int(3 * WL / c0 / cfg.time_step_duration)
    amps = [np.abs(st[1].detector_states[f"d{z}"]["fields"][-n_last:, 0, 0, 0, 0]).max() for z in (80, 100, 120)]
    omega = 2 * np.pi * c0 / WL
    from fdtdx_b200.constants import eps0, mu0
    k = omega * np.sqrt(mu0 * eps0 * (1 + 1j * sigma / (omega * eps0)))
    alpha = abs(k.imag)
    for a0, a1 in zip(amps[:-1], amps[1:]):
        assert abs(np.log(a0 / a1) / (20 * DX) - alpha) / alpha < 0.1


def test_lorentz_medium_phase_velocity():
    """ADE Lorentz pole (update.py:316-350; coefficient formulas dispersion.py:852-856): the phase
    velocity inside the medium matches n(w) = sqrt(eps_inf + d_eps w0^2 / (w0^2 - w^2 - i g w))."""
    nz = 150
    eps_inf, d_eps = 2.0, 1.5
    omega = 2 * np.pi * c0 / WL
    w0, gamma = 2.5 * omega, 0.0
    cfg_probe = fx.SimulationConfig(time=1e-15, grid=fx.UniformGrid(spacing=DX))
    dt = cfg_probe.time_step_duration
    # Lorentz ADE in the reference's normalisation: P'' + g P' + w0^2 P = d_eps w0^2 E
    denom = 1 + gamma * dt / 2
    c1 = (2 - (w0 * dt) ** 2) / denom
    c2 = -(1 - gamma * dt / 2) / denom
    c3 = d_eps * (w0 * dt) ** 2 / denom
    shape = (3, 3, nz)
    full = lambda v: np.full((1, 1, *shape), v, F)
    disp = {"c1": full(c1), "c2": full(c2), "c3": full(c3), "c4": None}
    objects, arrays, cfg = _line_scene(nz, eps_fn=lambda z: np.full(z.shape, eps_inf), time=70e-15, dispersive=disp,
                                       detectors=[_point("a", 60, ("Ex",)), _point("b", 66, ("Ex",))])
    st = yee.checkpointed_fdtd(arrays, objects, cfg)
    n_last = int(4 * WL / c0 / cfg.time_step_duration)
    a = st[1].detector_states["a"]["fields"][-n_last:, 0, 0, 0, 0]
    b = st[1].detector_states["b"]["fields"][-n_last:, 0, 0, 0, 0]
    dphi = np.angle(_fit_phase(b, dt, omega) / _fit_phase(a, dt, omega))
    n_meas = abs(dphi) / (6 * DX) / (omega / c0)
    n_expected = math.sqrt(eps_inf + d_eps * w0**2 / (w0**2 - omega**2))
    assert abs(n_meas - n_expected) / n_expected < 0.05
    assert np.isfinite(st[1].fields.dispersive_P_curr).all()


@pytest.mark.parametrize("kind", ["periodic", "pec", "pmc", "pml"])
def test_one_step_time_reversal(kind):
    """fdtd/test_time_reversal.py:251-327: forward then backward reconstructs E and H (atol 1e-5)."""
    shape = (8, 8, 8)
    rec = fx.Recorder(modules=[])
    cfg = fx.SimulationConfig(time=2e-15, grid=fx.UniformGrid(spacing=DX), gradient_config=fx.GradientConfig(recorder=rec))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = fx.boundary_objects_from_config(shape, cfg, kind, thickness=2)
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl], cfg, inv_permittivities=np.full((1, *shape), 0.5, F))
    rng = np.random.default_rng(0)
    arrays.fields.E[...] = (1e-3 * rng.standard_normal(arrays.fields.E.shape)).astype(F)
    arrays.fields.H[...] = (1e-3 * rng.standard_normal(arrays.fields.H.shape)).astype(F)
    for b in objects.boundary_objects:
        if isinstance(b, fx.PerfectElectricConductor):
            for comp in b.tangential_components:
                arrays.fields.E[(comp, *b.grid_slice)] = 0
        if isinstance(b, fx.PerfectMagneticConductor):
            for comp in b.tangential_components:
                arrays.fields.H[(comp, *b.grid_slice)] = 0
    E0, H0 = arrays.fields.E.copy(), arrays.fields.H.copy()
    st = yee.forward((0, arrays), cfg, objects, None, False, True, True)
    st = yee.backward(st, cfg, objects, None, False, False)
    assert st[0] == 0
    inner = (slice(None), slice(2, -2), slice(2, -2), slice(2, -2)) if kind == "pml" else (slice(None),) * 4
    assert np.allclose(st[1].fields.E[inner], E0[inner], atol=1e-5)
    assert np.allclose(st[1].fields.H[inner], H0[inner], atol=1e-5)


def _tensor_line_scene(nz, theta, eps_o, eps_e, pol, zstart=20, time=60e-15, detectors=()):
    """Periodic-xy line of cells, +z plane wave launched in vacuum, then a uniaxial medium whose principal
    axes are rotated by ``theta`` about z: eps = R diag(eps_o, eps_e, 1) R^T (all nine components stored)."""
    shape = (3, 3, nz)
    cfg = fx.SimulationConfig(time=time, grid=fx.UniformGrid(spacing=DX))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    types = {"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}
    bl = fx.boundary_objects_from_config(shape, cfg, types, thickness=10)
    c, s = math.cos(theta), math.sin(theta)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    inv_med = R @ np.diag([1.0 / eps_o, 1.0 / eps_e, 1.0]) @ R.T
    inv_eps = np.zeros((9, *shape), F)
    for i in range(3):
        for j in range(3):
            inv_eps[3 * i + j, :, :, :zstart] = 1.0 if i == j else 0.0
            inv_eps[3 * i + j, :, :, zstart:] = inv_med[i, j]
    src = fx.make_plane_source("s", ((0, 3), (0, 3), (14, 15)), cfg, inv_eps, direction="+", wave_character=fx.WaveCharacter(wavelength=WL),
                               fixed_E_polarization_vector=pol, normalize_by_energy=False)
    objs = [vol, *bl, src, *[d.place_on_grid(cfg) for d in detectors]]
    objects, arrays, _, cfg, _ = fx.place_objects(objs, cfg, inv_permittivities=inv_eps)
    return objects, arrays, cfg


@pytest.mark.parametrize("axis", ["ordinary", "extraordinary"])
def test_rotated_birefringence_full_tensor(axis):
    """tests/simulation/physics/test_birefringence.py re-posed for the nine-component tier (update.py:356-492):
    in a uniaxial medium whose axes are rotated by 30 degrees in the xy plane, a wave polarised along a
    principal axis is an eigen-polarisation - it keeps its polarisation and travels with that axis' index."""
    theta = math.radians(30.0)
    eps_o, eps_e = 2.25, 4.0
    e_o = (math.cos(theta), math.sin(theta), 0.0)
    e_e = (-math.sin(theta), math.cos(theta), 0.0)
    par, perp, n_med = (e_o, e_e, math.sqrt(eps_o)) if axis == "ordinary" else (e_e, e_o, math.sqrt(eps_e))
    dets = [_point("a", 60, comps=("Ex", "Ey")), _point("b", 66, comps=("Ex", "Ey"))]
    objects, arrays, cfg = _tensor_line_scene(140, theta, eps_o, eps_e, par, detectors=dets)
    st = yee.checkpointed_fdtd(arrays, objects, cfg)
    a = st[1].detector_states["a"]["fields"][:, :, 0, 0, 0]
    b = st[1].detector_states["b"]["fields"][:, :, 0, 0, 0]
    period_steps = WL / c0 / cfg.time_step_duration
    n_last = int(4 * period_steps)
    omega = 2 * np.pi * c0 / WL
    proj = lambda f, e: f[-n_last:, 0] * e[0] + f[-n_last:, 1] * e[1]
    pa = _fit_phase(proj(a, par), cfg.time_step_duration, omega)
    pb = _fit_phase(proj(b, par), cfg.time_step_duration, omega)
    k_meas = abs(np.angle(pb / pa)) / (6 * DX)
    k_expected = 2 * np.pi * n_med / WL
    assert abs(k_meas - k_expected) / k_expected < 0.05, (k_meas, k_expected)
    # eigen-polarisation: the orthogonal in-plane component stays small
    leak = np.abs(proj(a, perp)).max() / np.abs(proj(a, par)).max()
    assert leak < 0.05, leak


def test_stretched_grid_keeps_the_phase_velocity():
    """tests/simulation/physics/test_nonuniform_grid.py: on a mildly stretched z grid (+-8 % cell widths,
    metric scales of curl.py:10-39) the wave number measured over the PHYSICAL detector distance is
    still n k0."""
    nz, n_med = 140, 1.5
    shape = (3, 3, nz)
    cells = np.arange(nz, dtype=float)
    widths = DX * (1.0 + 0.08 * np.sin(2.0 * np.pi * (cells + 0.5) / nz))
    widths *= nz * DX / widths.sum()
    z_edges = np.concatenate([[0.0], np.cumsum(widths)])
    xy_edges = np.linspace(0.0, 3 * DX, 4)
    cfg = fx.SimulationConfig(time=60e-15, grid=fx.RectilinearGrid(xy_edges, xy_edges, z_edges))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    types = {"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}
    bl = fx.boundary_objects_from_config(shape, cfg, types, thickness=10)
    inv_eps = np.full((1, *shape), 1.0 / n_med**2, F)
    src = fx.make_plane_source("s", ((0, 3), (0, 3), (14, 15)), cfg, inv_eps, direction="+", wave_character=fx.WaveCharacter(wavelength=WL), normalize_by_energy=False)
    za, zb = 60, 70
    dets = [_point("a", za), _point("b", zb)]
    objs = [vol, *bl, src, *[d.place_on_grid(cfg) for d in dets]]
    objects, arrays, _, cfg, _ = fx.place_objects(objs, cfg, inv_permittivities=inv_eps)
    st = yee.checkpointed_fdtd(arrays, objects, cfg)
    a = st[1].detector_states["a"]["fields"][:, :, 0, 0, 0]
    b = st[1].detector_states["b"]["fields"][:, :, 0, 0, 0]
    period_steps = WL / c0 / cfg.time_step_duration
    n_last = int(4 * period_steps)
    omega = 2 * np.pi * c0 / WL
    pa = _fit_phase(a[-n_last:, 0], cfg.time_step_duration, omega)
    pb = _fit_phase(b[-n_last:, 0], cfg.time_step_duration, omega)
    centers = 0.5 * (z_edges[:-1] + z_edges[1:])
    dist = centers[zb] - centers[za]
    k_meas = abs(np.angle(pb / pa)) / dist
    k_expected = 2 * np.pi * n_med / WL
    assert abs(dist - (zb - za) * DX) / ((zb - za) * DX) > 0.01  # the detectors really sit on stretched cells
    assert abs(k_meas - k_expected) / k_expected < 0.05, (k_meas, k_expected)


def test_drude_medium_phase_velocity():
    """ADE Drude pole (omega_0 = 0, coupling_sq = omega_p^2; coefficient formulas dispersion.py:852-856):
    below-plasma-frequency dielectric response n(w) = sqrt(eps_inf - w_p^2 / (w^2 + i g w)), lossless here."""
    nz = 150
    eps_inf = 4.0
    omega = 2 * np.pi * c0 / WL
    wp, gamma = 1.2 * omega, 0.0
    dt = fx.SimulationConfig(time=1e-15, grid=fx.UniformGrid(spacing=DX)).time_step_duration
    denom = 1 + gamma * dt / 2
    c1 = 2.0 / denom
    c2 = -(1 - gamma * dt / 2) / denom
    c3 = (wp * dt) ** 2 / denom
    shape = (3, 3, nz)
    full = lambda v: np.full((1, 1, *shape), v, F)
    disp = {"c1": full(c1), "c2": full(c2), "c3": full(c3), "c4": None}
    objects, arrays, cfg = _line_scene(nz, eps_fn=lambda z: np.full(z.shape, eps_inf), time=70e-15, dispersive=disp,
                                       detectors=[_point("a", 60, ("Ex",)), _point("b", 66, ("Ex",))])
    st = yee.checkpointed_fdtd(arrays, objects, cfg)
    n_last = int(4 * WL / c0 / cfg.time_step_duration)
    a = st[1].detector_states["a"]["fields"][-n_last:, 0, 0, 0, 0]
    b = st[1].detector_states["b"]["fields"][-n_last:, 0, 0, 0, 0]
    dphi = np.angle(_fit_phase(b, dt, omega) / _fit_phase(a, dt, omega))
    n_meas = abs(dphi) / (6 * DX) / (omega / c0)
    n_expected = math.sqrt(eps_inf - wp**2 / omega**2)
    assert abs(n_meas - n_expected) / n_expected < 0.05, (n_meas, n_expected)
