"""Pins the CPU oracle with the reference's own closed-form unit tests, re-typed against
``oracle/yee.py``.  Each test names the reference test it restates
(paths relative to ``/root/reference/tests/unit``).  The reference ships no golden field vectors;
these known-answer formulas are what pins the hot path there, and therefore here.
"""

import math
from types import SimpleNamespace

import numpy as np
import pytest

import fdtdx_b200 as fx
from fdtdx_b200.constants import eta0
from fdtdx_b200.container import ArrayContainer, FieldState
from oracle import yee

F = np.float32


def _cfg(c=None):
    cfg = fx.SimulationConfig(time=400e-15, grid=fx.UniformGrid(spacing=1.0), courant_factor=0.99)
    return cfg


def _nonuniform_cfg():
    grid = fx.RectilinearGrid([0.0, 1.0, 3.0, 6.0, 10.0], [0.0, 1.0, 2.5, 5.0, 9.0], [0.0, 1.0, 4.0, 8.0, 13.0])
    return fx.SimulationConfig(time=400e-15, grid=grid, courant_factor=0.99)


def _objects(pml=(), boundaries=(), shape=(4, 4, 4)):
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    return fx.ObjectContainer([vol, *boundaries, *pml])


# ---------------------------------------------------------------- curl (core/physics/test_curl.py)
def test_curl_E_linear_field():
    """test_curl.py:162-183: E_x = y, E_y = -x gives curl_z = -2."""
    n = 6
    X, Y, _ = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    E = np.stack([Y, -X, np.zeros_like(X)]).astype(F)
    curl, psi = yee.curl_E(_cfg(), yee.pad_fields(E, (True, True, True)), {}, _objects(), True)
    assert curl.shape == (3, 6, 6, 6)
    assert np.allclose(curl[2][:-1, :-1], -2.0, atol=0.1)
    assert not psi


def test_curl_H_linear_field():
    """test_curl.py:339-360: H_x = z, H_z = -x gives curl_y = 2."""
    n = 6
    X, _, Z = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    H = np.stack([Z, np.zeros_like(X), -X]).astype(F)
    curl, _ = yee.curl_H(_cfg(), yee.pad_fields(H, (True, True, True)), {}, _objects(), True)
    assert np.allclose(curl[1][1:, :, 1:], 2.0, atol=0.1)


def test_curl_zero_field_nonperiodic():
    """test_curl.py:186-201."""
    E = np.zeros((3, 4, 4, 4), F)
    curl, psi = yee.curl_E(_cfg(), yee.pad_fields(E, (False, False, False)), {}, _objects(), True)
    assert np.allclose(curl, 0.0) and not psi


@pytest.mark.parametrize("is_E", [True, False])
def test_curl_pml_scatter_signs(is_E, monkeypatch):
    """test_curl.py:204-255 / 381-432: an x-axis PML (a=0) subtracts corr_1 from component i=1 (y)
    and adds corr_2 to component j=2 (z), and the updated psi is passed through."""
    n = 5
    sl = ((0, n), (0, n), (1, 4))
    pml = fx.PerfectlyMatchedLayer(name="pml_x", grid_slice_tuple=sl, axis=0, direction="-")
    calls = {}

    def fake_step(p, d1, d2, p1, p2, is_curl_E, simulate):
        calls["args"] = (is_curl_E, simulate, p1, p2)
        o = np.ones((n, n, 3), F)
        return o * F(0.1), o * F(0.2), o * F(0.3), o * F(0.4)

    monkeypatch.setattr(yee, "step_cpml", fake_step)
    psi0 = (np.zeros((n, n, 3), F), np.zeros((n, n, 3), F))
    Fz = np.zeros((3, n, n, n), F)
    fn = yee.curl_E if is_E else yee.curl_H
    curl, psi = fn(_cfg(), yee.pad_fields(Fz, (False,) * 3), {"pml_x": psi0}, _objects(pml=[pml], shape=(n, n, n)), False)
    assert calls["args"][0] is is_E and calls["args"][1] is False
    assert np.allclose(psi["pml_x"][0], 0.3) and np.allclose(psi["pml_x"][1], 0.4)
    gs = pml.grid_slice
    assert np.allclose(curl[1][gs], -0.1) and np.allclose(curl[2][gs], 0.2)
    assert np.allclose(curl[0], 0.0)


def test_curl_E_nonuniform_metric_linear_field_exact():
    """test_curl.py:279-300: local metric factors recover the physical curl exactly (1e-6)."""
    cfg = _nonuniform_cfg()
    g = cfg.grid
    X, Y, _ = np.meshgrid(g.x_edges[:-1], g.y_edges[:-1], g.z_edges[:-1], indexing="ij")
    E = np.stack([Y, -X, np.zeros_like(X)]).astype(F)
    curl, _ = yee.curl_E(cfg, yee.pad_fields(E, (False,) * 3), {}, _objects(), False)
    assert np.allclose(curl[2][:-1, :-1, :], -2.0, atol=1e-6)


def test_curl_E_nonuniform_quadratic_field():
    """test_curl.py:303-336: forward derivatives use the local stretched widths."""
    cfg = _nonuniform_cfg()
    g = cfg.grid
    y, z = g.y_edges[:-1], g.z_edges[:-1]
    _, Y, Z = np.meshgrid(g.x_edges[:-1], y, z, indexing="ij")
    E = np.stack([np.zeros_like(Y), Z**2, Y**2]).astype(F)
    curl, _ = yee.curl_E(cfg, yee.pad_fields(E, (False,) * 3), {}, _objects(), False)
    expected = (y[:-1] + y[1:])[None, :, None] - (z[:-1] + z[1:])[None, None, :]
    assert np.allclose(curl[0][:, :-1, :-1], expected, atol=1e-6)


def test_curl_H_nonuniform_metric_backward_average_widths():
    """test_curl.py:454-506: the backward stencil divides by the mean of adjacent cell widths."""
    cfg = _nonuniform_cfg()
    g = cfg.grid
    xc = g.centers(0)
    zc = g.centers(2)
    X, _, Z = np.meshgrid(xc, g.centers(1), zc, indexing="ij")
    H = np.stack([Z, np.zeros_like(X), -X]).astype(F)  # sampled at cell centres: dz Hx = 1, dx Hz = -1
    curl, _ = yee.curl_H(cfg, yee.pad_fields(H, (False,) * 3), {}, _objects(), False)
    assert np.allclose(curl[1][1:, :, 1:], 2.0, atol=1e-5)


def test_interpolate_fields_basic_and_region_parity():
    """test_curl.py:44-157: uniform fields, and region-restricted == full-domain + slice, on uniform
    and stretched grids (tests/integration/fdtd/test_detector_region_parity.py)."""
    E = np.ones((3, 5, 5, 5), F)
    H = np.ones((3, 5, 5, 5), F) * F(0.5)
    Ei, Hi = yee.interpolate_fields(yee.pad_fields(E, (False,) * 3), yee.pad_fields(H, (False,) * 3))
    assert Ei.shape == (3, 5, 5, 5) and np.all(Ei[2] == 1.0)
    assert np.allclose(Ei[0][1:, :, :-1], 1.0) and np.allclose(Hi[2][1:, 1:, :-1], 0.5)
    rng = np.random.default_rng(0)
    for cfg in (None, _nonuniform_cfg()):
        n = 4
        E = rng.standard_normal((3, n, n, n)).astype(F)
        H = rng.standard_normal((3, n, n, n)).astype(F)
        fullE, fullH = yee.interpolate_fields(yee.pad_fields(E, (False,) * 3), yee.pad_fields(H, (False,) * 3), config=cfg)
        gst = ((1, 3), (1, 3), (1, 3))
        block = (slice(None), *(slice(s - 1, e + 1) for s, e in gst))
        region = (slice(None), *(slice(s, e) for s, e in gst))
        rE, rH = yee.interpolate_fields(E[block], H[block], config=cfg, region_slice=gst)
        assert np.allclose(rE, fullE[region], atol=1e-6) and np.allclose(rH, fullH[region], atol=1e-6)


# ---------------------------------------------------------------- update algebra (fdtd/test_update.py)
def _arrays(shape=(4, 4, 4), E=None, H=None, inv_eps=None, sigma_E=None, sigma_H=None, inv_mu=1.0):
    z = lambda: np.zeros((3, *shape), F)
    return ArrayContainer(
        fields=FieldState(E=z() if E is None else E, H=z() if H is None else H, psi_E={}, psi_H={}),
        inv_permittivities=np.ones((1, *shape), F) if inv_eps is None else inv_eps,
        inv_permeabilities=inv_mu,
        detector_states={},
        recording_state=None,
        electric_conductivity=sigma_E,
        magnetic_conductivity=sigma_H,
    )


def _patched_cfg(c):
    cfg = _cfg()
    object.__setattr__(cfg, "courant_factor", c * math.sqrt(3))
    return cfg


def test_update_E_formula_lossless_and_lossy(monkeypatch):
    """test_update.py:288-314: E' = E + c*curl*inv_eps ; lossy: (1-h)E/(1+h), h = c*sigma*eta0*inv_eps/2."""
    shape = (4, 4, 4)
    c = 0.5
    curl = np.ones((3, *shape), F) * F(2.0)
    monkeypatch.setattr(yee, "curl_H", lambda *a, **k: (curl, {}))
    out = yee.update_E(0, _arrays(E=np.ones((3, *shape), F)), _objects(), _patched_cfg(c), False)
    assert np.allclose(out.fields.E, 1.0 + c * 2.0 * 1.0)
    monkeypatch.setattr(yee, "curl_H", lambda *a, **k: (np.zeros((3, *shape), F), {}))
    sig = np.ones((1, *shape), F) * F(1e-4)
    out = yee.update_E(0, _arrays(E=np.ones((3, *shape), F) * 2, sigma_E=sig), _objects(), _patched_cfg(c), False)
    half = c * 1e-4 * eta0 / 2
    assert np.allclose(out.fields.E, (1 - half) * 2.0 / (1 + half), rtol=1e-5)


def test_update_H_formula_lossless_and_lossy(monkeypatch):
    """test_update.py:604-630: H' = H - c*curl*inv_mu ; lossy with h = c*sigma_H/eta0*inv_mu/2."""
    shape = (4, 4, 4)
    c = 0.5
    monkeypatch.setattr(yee, "curl_E", lambda *a, **k: (np.ones((3, *shape), F) * F(2.0), {}))
    out = yee.update_H(0, _arrays(H=np.ones((3, *shape), F)), _objects(), _patched_cfg(c), False)
    assert np.allclose(out.fields.H, 1.0 - c * 2.0)
    monkeypatch.setattr(yee, "curl_E", lambda *a, **k: (np.zeros((3, *shape), F), {}))
    sig = np.ones((1, *shape), F) * F(50.0)
    out = yee.update_H(0, _arrays(H=np.ones((3, *shape), F) * 2, sigma_H=sig), _objects(), _patched_cfg(c), False)
    half = c * 50.0 / eta0 / 2
    assert np.allclose(out.fields.H, (1 - half) * 2.0 / (1 + half), rtol=1e-5)


@pytest.mark.parametrize("lossy", [False, True])
def test_reverse_undoes_forward(lossy):
    """test_update.py:408-435 / 484-510: update_X_reverse(update_X(state)) == state."""
    rng = np.random.default_rng(1)
    shape = (5, 4, 6)
    cfg = fx.SimulationConfig(time=1e-13, grid=fx.UniformGrid(spacing=50e-9))
    bl = fx.boundary_objects_from_config(shape, cfg, "periodic")
    obj = _objects(boundaries=bl, shape=shape)
    E = rng.standard_normal((3, *shape)).astype(F)
    H = rng.standard_normal((3, *shape)).astype(F)
    kw = {}
    if lossy:
        kw = dict(sigma_E=(1e-4 * rng.random((1, *shape))).astype(F), sigma_H=(20 * rng.random((1, *shape))).astype(F))
    arr = _arrays(shape, E=E.copy(), H=H.copy(), inv_eps=(1 / (1 + rng.random((3, *shape)))).astype(F), **kw)
    a1 = yee.update_E(0, arr, obj, cfg, True)
    a2 = yee.update_E_reverse(0, a1, obj, cfg)
    assert np.allclose(a2.fields.E, E, atol=2e-6)
    b1 = yee.update_H(0, arr, obj, cfg, True)
    b2 = yee.update_H_reverse(0, b1, obj, cfg)
    assert np.allclose(b2.fields.H, H, atol=2e-6)


def test_halo_rules():
    """test_update.py:190-208: wrap halo on periodic axes, zero halo elsewhere."""
    f = np.arange(3 * 2 * 2 * 2, dtype=F).reshape(3, 2, 2, 2) + 1
    p = yee.pad_fields(f, (True, False, False))
    assert p.shape == (3, 4, 4, 4)
    assert np.array_equal(p[:, 0, 1:-1, 1:-1], f[:, -1]) and np.array_equal(p[:, -1, 1:-1, 1:-1], f[:, 0])
    assert np.all(p[:, :, 0] == 0) and np.all(p[:, :, :, -1] == 0)


def test_dispersive_reverse_raises():
    shape = (3, 3, 3)
    arr = _arrays(shape)
    arr = arr.aset("fields->dispersive_P_curr", np.zeros((1, 3, *shape), F))
    with pytest.raises(NotImplementedError):
        yee.update_E_reverse(0, arr, _objects(shape=shape), _cfg())


# ---------------------------------------------------------------- fdtd/test_fdtd_misc.py
def test_anisotropic_update_matrices():
    """test_fdtd_misc.py:158-250."""
    sp = (5, 5, 5)
    inv = np.arange(3 * 3 * 125, dtype=F).reshape(3, 3, *sp)
    A, B = yee.compute_anisotropic_update_matrices(inv, None, 0.5, 1.0)
    eye = np.eye(3)[:, :, None, None, None] * np.ones((1, 1, *sp))
    assert np.allclose(A, eye, atol=1e-6) and np.allclose(B, 0.5 * inv, atol=1e-6)
    Ar, Br = yee.compute_anisotropic_update_matrices(inv, None, 0.5, 1.0, reverse=True)
    assert np.allclose(Ar, eye, atol=1e-6) and np.allclose(Br, 0.5 * inv, atol=1e-6)
    inv = (np.eye(3)[:, :, None, None, None] * np.ones((1, 1, 3, 3, 3))).astype(F)
    sig = (inv * 0.5).astype(F)
    A, _ = yee.compute_anisotropic_update_matrices(inv, sig, 0.5, 1.0)
    Ar, _ = yee.compute_anisotropic_update_matrices(inv, sig, 0.5, 1.0, reverse=True)
    f = 0.5 * 1.0 / 2 * 0.5
    assert np.allclose(A[0, 0], (1 - f) / (1 + f), atol=1e-6) and np.allclose(Ar[0, 0], (1 + f) / (1 - f), atol=1e-6)
    assert not np.allclose(A, Ar)


def test_avg_anisotropic_components():
    """test_fdtd_misc.py:257-403: 4-point means at the other component's Yee location."""
    rng = np.random.default_rng(2)
    f = rng.standard_normal((3, 5, 5, 5)).astype(F)
    p = f  # treated as an already padded array; the result covers the interior 3x3x3
    out = yee.avg_anisotropic_E_component(p, component=0, location=1)
    i = np.s_[1:-1, 1:-1, 1:-1]
    exp = (p[0][1:-1, 1:-1, 1:-1] + p[0][1:-1, 2:, 1:-1] + p[0][:-2, 1:-1, 1:-1] + p[0][:-2, 2:, 1:-1]) / 4
    assert np.allclose(out, exp, atol=1e-6)
    out = yee.avg_anisotropic_H_component(p, component=2, location=0)
    exp = (p[2][1:-1, 1:-1, 1:-1] + p[2][:-2, 1:-1, 1:-1] + p[2][1:-1, 1:-1, 2:] + p[2][:-2, 1:-1, 2:]) / 4
    assert np.allclose(out, exp, atol=1e-6)
    const = np.ones((3, 5, 5, 5), F) * 3
    w = tuple(np.array([1, 1, 2, 3, 3], F).reshape([5 if a == ax else 1 for a in range(3)]) for ax in range(3))
    assert np.allclose(yee.avg_anisotropic_E_component(const, 0, 2, w), 3.0)
    assert np.allclose(yee.avg_anisotropic_H_component(const, 1, 0, w), 3.0)


def test_interface_gather_scatter_roundtrip():
    """test_fdtd_misc.py:48-150 + interfaces/test_recorder.py: collect then add restores the planes."""
    shape = (8, 7, 9)
    rec = fx.Recorder(modules=[])
    cfg = fx.SimulationConfig(time=2e-15, grid=fx.UniformGrid(spacing=50e-9), gradient_config=fx.GradientConfig(recorder=rec))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = fx.boundary_objects_from_config(shape, cfg, "pml", thickness=2)
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl], cfg, inv_permittivities=np.ones((1, *shape), F))
    rng = np.random.default_rng(3)
    arrays.fields.E[...] = rng.standard_normal(arrays.fields.E.shape).astype(F)
    arrays.fields.H[...] = rng.standard_normal(arrays.fields.H.shape).astype(F)
    E0, H0 = arrays.fields.E.copy(), arrays.fields.H.copy()
    arrays = yee.collect_interfaces(1, arrays, objects, cfg)
    tampered = arrays.aset("fields->E", np.zeros_like(E0)).aset("fields->H", np.zeros_like(H0))
    restored = yee.add_interfaces(1, tampered, objects, cfg)
    for pml in objects.pml_objects:
        sl = (slice(None), *pml.interface_slice())
        assert np.array_equal(restored.fields.E[sl], E0[sl]) and np.array_equal(restored.fields.H[sl], H0[sl])
    lo = bl[0]
    assert lo.interface_slice_tuple()[0] == (1, 2) and bl[1].interface_slice_tuple()[0] == (shape[0] - 2, shape[0] - 1)


# ---------------------------------------------------------------- objects/boundaries/test_perfectly_matched_layer.py
def test_pml_profiles_and_defaults():
    """test_perfectly_matched_layer.py:193-300: defaults, grading monotonic towards the outer edge,
    b in (0, 1], a <= 0, interface cell of the E profile has zero depth."""
    cfg = fx.SimulationConfig(time=1e-13, grid=fx.UniformGrid(spacing=50e-9))
    for direction, sl in (("-", ((0, 10), (0, 4), (0, 4))), ("+", ((30, 40), (0, 4), (0, 4)))):
        p = fx.PerfectlyMatchedLayer(name="p", grid_slice_tuple=sl, axis=0, direction=direction).place_on_grid(cfg)
        assert p.kappa_start == 1.0 and p.kappa_end == 1.0 and p.sigma_order == 3.0 and p.alpha_end == 0.0
        bE, aE = p.pml_b_E.reshape(-1), p.pml_a_E.reshape(-1)
        assert np.all((bE > 0) & (bE <= 1)) and np.all(aE <= 0) and np.all(np.isfinite(aE))
        inner_to_outer = bE if direction == "+" else bE[::-1]
        assert np.all(np.diff(inner_to_outer) <= 1e-7)  # absorption grows towards the outer boundary
        sigma_end = -(3 + 1) * math.log(1e-6) / (2 * eta0 * 10 * 50e-9)
        assert abs(p.sigma_end - sigma_end) / sigma_end < 1e-5
        assert np.allclose(p.inv_kappa_E, 1.0) and p.pml_a_E.shape == (10, 1, 1)


def test_step_cpml_formulas():
    """perfectly_matched_layer.py:138-190 (the reference mocks this in test_curl.py)."""
    cfg = fx.SimulationConfig(time=1e-13, grid=fx.UniformGrid(spacing=50e-9))
    p = fx.PerfectlyMatchedLayer(name="p", grid_slice_tuple=((0, 4), (0, 3), (0, 3)), axis=0, direction="-").place_on_grid(cfg)
    rng = np.random.default_rng(4)
    d1, d2, s1, s2 = (rng.standard_normal((4, 3, 3)).astype(F) for _ in range(4))
    c1, c2, n1, n2 = yee.step_cpml(p, d1, d2, s1, s2, True, True)
    assert np.allclose(n1, p.pml_b_H * s1 + p.pml_a_H * d1) and np.array_equal(c1, n1) and np.array_equal(c2, n2)
    c1, c2, n1, n2 = yee.step_cpml(p, d1, d2, s1, s2, False, False)
    assert np.array_equal(n1, s1) and np.array_equal(c2, s2)
    p.kappa_end = 3.0
    p.place_on_grid(cfg)
    c1, _, n1, _ = yee.step_cpml(p, d1, d2, s1, s2, False, True)
    assert np.allclose(c1, (p.inv_kappa_E - 1) * d1 + n1)


# ---------------------------------------------------------------- objects/sources/test_tfsf.py
def _plane_scene(eps_tier=1):
    shape = (8, 8, 10)
    cfg = fx.SimulationConfig(time=20e-15, grid=fx.UniformGrid(spacing=50e-9))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = fx.boundary_objects_from_config(shape, cfg, "periodic")
    inv_eps = np.full((eps_tier, *shape), 0.5, F)
    src = fx.make_plane_source("s", ((0, 8), (0, 8), (4, 5)), cfg, inv_eps, direction="+", wave_character=fx.WaveCharacter(wavelength=1e-6))
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl, src], cfg, inv_permittivities=inv_eps)
    return objects, arrays, cfg, src


@pytest.mark.parametrize("tier", [1, 3])
def test_tfsf_injection_only_changes_plane_and_inverse_undoes(tier):
    """test_tfsf.py:304-420, 452-500."""
    objects, arrays, cfg, src = _plane_scene(tier)
    t = 37.0
    E0 = np.zeros_like(arrays.fields.E)
    E1 = yee.tfsf_update_E(src, E0, arrays.inv_permittivities, F(t), False, cfg)
    changed = np.abs(E1) > 0
    assert changed.any() and not changed[:, :, :, :4].any() and not changed[:, :, :, 5:].any()
    assert not changed[2].any()  # the face-normal component is untouched for diagonal media
    E2 = yee.tfsf_update_E(src, E1, arrays.inv_permittivities, F(t), True, cfg)
    assert np.allclose(E2, E0, atol=1e-9)
    H1 = yee.tfsf_update_H(src, E0.copy(), 1.0, F(t + 0.5), False, cfg)
    assert (np.abs(H1[:, :, :, 4]) > 0).any() and not (np.abs(H1[2]) > 0).any()
    H2 = yee.tfsf_update_H(src, H1, 1.0, F(t + 0.5), True, cfg)
    assert np.allclose(H2, 0.0, atol=1e-9)
    neg = fx.make_plane_source("n", ((0, 8), (0, 8), (4, 5)), cfg, arrays.inv_permittivities, direction="-", wave_character=fx.WaveCharacter(wavelength=1e-6))
    En = yee.tfsf_update_E(neg, E0, arrays.inv_permittivities, F(t), False, cfg)
    assert np.abs(En).max() > 0


def test_temporal_profiles():
    """profile.py:263-273, 322-345, 412-439."""
    period = 5e-15
    cw = fx.SingleFrequencyProfile()
    t = np.array([0.0, 0.25, 2.0, 4.0, 10.0], F) * F(period)
    a = cw.get_amplitude(t, period)
    exp = np.cos(2 * np.pi * t.astype(np.float64) / period + np.pi) * np.clip(t / (4 * period), 0, 1)
    assert np.allclose(a, exp, atol=2e-5)
    gp = fx.GaussianPulseProfile(spectral_width=fx.WaveCharacter(frequency=4e13), center_wave=fx.WaveCharacter(frequency=2e14))
    sig = 1 / (2 * np.pi * 4e13)
    a = gp.get_amplitude(np.array([6 * sig], F), 0.0)
    assert abs(float(a[0]) - math.cos(2 * np.pi * 2e14 * 6 * sig)) < 1e-3
    tab = fx.CustomTimeSignalProfile(signal=np.array([0.0, 1.0, 3.0], F), time_step_duration=1.0)
    assert np.allclose(tab.get_amplitude(np.array([0.5, 1.5, 5.0, -1.0], F), 0.0), [0.5, 2.0, 0.0, 0.0])


# ---------------------------------------------------------------- detectors
def test_phasor_single_frequency_accumulation():
    """objects/detectors/test_phasor.py:99-316: state += EH * exp(i w t dt) * 2/N."""
    cfg = fx.SimulationConfig(time=40e-15, grid=fx.UniformGrid(spacing=50e-9))
    wc = fx.WaveCharacter(wavelength=1e-6)
    det = fx.PhasorDetector(name="p", grid_slice_tuple=((0, 2), (0, 2), (0, 2)), wave_characters=(wc,), components=("Ex",)).place_on_grid(cfg)
    T = cfg.time_steps_total
    state = det.init_state()
    assert state["phasor"].shape == (1, 1, 1, 2, 2, 2) and state["phasor"].dtype == np.complex64
    w = 2 * np.pi * wc.get_frequency()
    amp, phase = 0.7, 0.3
    for t in range(T):
        E = np.full((3, 2, 2, 2), amp * np.cos(w * t * cfg.time_step_duration + phase), F)
        state = yee.detector_update(det, t, E, E, state, None, 1.0)
    got = state["phasor"][0, 0, 0, 0, 0, 0]
    assert abs(abs(got) - amp) / amp < 2e-2 and abs(np.angle(got) + phase) < 5e-2
    assert det._static_scale() == 2 / T


def test_phasor_pulse_stride_scaling():
    """objects/detectors/test_phasor_subsample.py:84-156."""
    cfg = fx.SimulationConfig(time=20e-15, grid=fx.UniformGrid(spacing=50e-9))
    wc = fx.WaveCharacter(wavelength=1e-6)
    d = fx.PhasorDetector(name="p", grid_slice_tuple=((0, 1),) * 3, wave_characters=(wc,), scaling_mode="pulse", dft_subsample=3).place_on_grid(cfg)
    assert d._static_scale() == 3 and int(d._is_on_at_time_step_arr.sum()) == math.ceil(cfg.time_steps_total / 3)
    assert d._is_on_at_time_step_arr[0] and not d._is_on_at_time_step_arr[1] and d._is_on_at_time_step_arr[3]


def test_energy_field_poynting_detectors():
    """objects/detectors/test_{energy,field,poynting_flux}.py + core/physics/metrics.py:55-117."""
    cfg = fx.SimulationConfig(time=5e-15, grid=fx.UniformGrid(spacing=50e-9))
    sl = ((0, 3), (0, 4), (0, 1))
    rng = np.random.default_rng(5)
    E = rng.standard_normal((3, 3, 4, 1)).astype(F)
    H = rng.standard_normal((3, 3, 4, 1)).astype(F)
    ie = (0.3 + rng.random((3, 3, 4, 1))).astype(F)
    en = yee.compute_energy(E, H, ie, 1.0)
    assert np.allclose(en, 0.5 * ((E**2 / ie).sum(0) + (H**2).sum(0)), rtol=1e-5)
    S = yee.compute_poynting_flux(E, H)
    assert np.allclose(S, np.cross(E, H, axis=0), rtol=1e-5)
    pf = fx.PoyntingFluxDetector(name="pf", grid_slice_tuple=sl, direction="-").place_on_grid(cfg)
    st = yee.detector_update(pf, 2, E, H, pf.init_state(), ie, 1.0)
    assert np.allclose(st["poynting_flux"][2, 0], -(S[2] * (50e-9) ** 2).sum(), rtol=1e-4)
    fd = fx.FieldDetector(name="f", grid_slice_tuple=sl, components=("Ey", "Hz"), reduce_volume=True).place_on_grid(cfg)
    st = yee.detector_update(fd, 1, E, H, fd.init_state(), ie, 1.0)
    assert np.allclose(st["fields"][1], [E[1].mean(), H[2].mean()], rtol=1e-4)
    ed = fx.EnergyDetector(name="e", grid_slice_tuple=sl, as_slices=True).place_on_grid(cfg)
    st = yee.detector_update(ed, 0, E, H, ed.init_state(), ie, 1.0)
    assert st["XY Plane"].shape[1:] == (3, 4) and np.allclose(st["XZ Plane"][0], en.mean(axis=1), rtol=1e-5)


# ---------------------------------------------------------------- interfaces/test_recorder.py, test_time_filter.py
def test_linear_reconstruct_every_k_tables():
    """test_time_filter.py:49-140, 205-230 and test_recorder.py:69-190: saved steps 0,K,..,T-1;
    unsaved steps interpolate linearly between the neighbouring saved slots."""
    rec = fx.Recorder(modules=[fx.LinearReconstructEveryK(k=5), fx.DtypeConversion(dtype="float8_e4m3fnuz")]).init_tables(23)
    saved = [t for t in range(23) if rec.slot_of_time[t] >= 0]
    assert saved == [0, 5, 10, 15, 20, 22] and rec._latent_array_size == 6 and rec.elem_bytes == 1
    assert rec.replay_a[7] == 1 and rec.replay_b[7] == 2 and abs(rec.replay_w[7] - 0.4) < 1e-7
    assert rec.replay_a[21] == 4 and rec.replay_b[21] == 5 and abs(rec.replay_w[21] - 0.5) < 1e-7
    assert rec.replay_a[10] == rec.replay_b[10] == 2 and rec.replay_w[10] == 0
    plain = fx.Recorder(modules=[]).init_tables(7)
    assert list(plain.slot_of_time) == list(range(7)) and plain.elem_bytes == 4


def test_recorder_dtype_roundtrip_and_interpolated_decompress():
    """test_recorder.py:241-290."""
    shape = (6, 6, 6)
    rec = fx.Recorder(modules=[fx.LinearReconstructEveryK(k=2), fx.DtypeConversion(dtype="bfloat16")])
    cfg = fx.SimulationConfig(time=1.2e-15, grid=fx.UniformGrid(spacing=50e-9), gradient_config=fx.GradientConfig(recorder=rec))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = fx.boundary_objects_from_config(shape, cfg, {"min_x": "pml", "max_x": "pml", "min_y": "periodic", "max_y": "periodic", "min_z": "periodic", "max_z": "periodic"}, thickness=2)
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl], cfg, inv_permittivities=np.ones((1, *shape), F))
    T = cfg.time_steps_total
    assert T >= 5
    for t in (0, 2):
        arrays.fields.E[...] = F(1.0 + t)
        arrays.fields.H[...] = F(-1.0 - t)
        arrays = yee.collect_interfaces(t, arrays, objects, cfg)
    arrays.fields.E[...] = 0
    out = yee.add_interfaces(1, arrays, objects, cfg)
    pml = objects.pml_objects[0]
    assert np.allclose(out.fields.E[(slice(None), *pml.interface_slice())], 2.0)  # midpoint of 1 and 3
    assert np.allclose(out.fields.H[(slice(None), *pml.interface_slice())], -2.0)
