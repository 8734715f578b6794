"""Independent pins of the host-side tables that the oracle SHARES with the product
(``oracle/yee.py`` imports ``fdtdx_b200.boundaries/detectors/sources/recorder``: a wrong table there
is invisible to every CUDA-vs-oracle comparison).  Each expectation below is typed directly from the
closed forms in the reference source - none of it calls into ``fdtdx_b200`` to build the expected value.

CPU only (``-m "not gpu"``).
"""

import math

import numpy as np
import pytest

import fdtdx_b200 as fx
from fdtdx_b200 import constants, dispersion
from fdtdx_b200.plan import metric_scales
from fdtdx_b200.sources import calculate_time_offset_yee

F = np.float32
C0 = 299792458.0


# ---------------------------------------------------------------------------------------------
# OnOffSwitch (core/switch.py:112-215)
# ---------------------------------------------------------------------------------------------
def test_switch_interval_window_and_fixed_lists():
    dt, T = 1e-16, 200
    # interval only: on at t % interval == 0 (switch.py:210-213)
    on = fx.OnOffSwitch(interval=3).calculate_on_list(T, dt)
    assert on == [t % 3 == 0 for t in range(T)]
    # start / end times are inclusive bounds on t*dt (switch.py:206-208)
    on = fx.OnOffSwitch(start_time=20 * dt, end_time=50 * dt).calculate_on_list(T, dt)
    exp = [(20 * dt <= t * dt) and (t * dt <= 50 * dt) for t in range(T)]
    assert on == exp and sum(on) in (30, 31)
    # periods: start_after_periods * period, on_for_periods * period
    per = 10 * dt
    on = fx.OnOffSwitch(start_after_periods=2.0, on_for_periods=3.0, period=per).calculate_on_list(T, dt)
    assert on == [(2.0 * per <= t * dt) and (t * dt <= 2.0 * per + 3.0 * per) for t in range(T)]
    # fixed list incl. a negative index (python indexing, used by optimize_ceviche_corner.py: [-1])
    on = fx.OnOffSwitch(fixed_on_time_steps=[3, 7, -1]).calculate_on_list(T, dt)
    assert [t for t, v in enumerate(on) if v] == [3, 7, T - 1]
    assert not any(fx.OnOffSwitch(is_always_off=True).calculate_on_list(T, dt))


def test_switch_time_step_to_array_index_is_a_running_count():
    dt, T = 1e-16, 50
    sw = fx.OnOffSwitch(interval=4, start_time=8 * dt)
    on = sw.calculate_on_list(T, dt)
    idx = sw.calculate_time_step_to_on_arr_idx(T, dt)
    count = 0
    for t in range(T):
        if on[t]:
            assert idx[t] == count
            count += 1
        else:
            assert idx[t] == -1
    assert count == sum(on) > 0


# ---------------------------------------------------------------------------------------------
# every-K recorder slot maps (interfaces/time_filter.py:166-190, 237-250)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,K", [(23, 5), (21, 5), (1311, 5), (17, 1), (9, 4)])
def test_every_k_slot_and_replay_tables(T, K):
    rec = fx.Recorder(modules=[fx.LinearReconstructEveryK(k=K)]).init_tables(T)
    saved = list(range(0, T, K))
    if saved[-1] != T - 1:
        saved.append(T - 1)
    assert rec._latent_array_size == len(saved)
    for t in range(T):
        if t in saved:
            s = saved.index(t)
            assert rec.slot_of_time[t] == s
            assert rec.replay_a[t] == s and rec.replay_b[t] == s
        else:
            assert rec.slot_of_time[t] == -1
            prev = max(x for x in saved if x < t)
            nxt = min(x for x in saved if x > t)
            assert rec.replay_a[t] == saved.index(prev) and rec.replay_b[t] == saved.index(nxt)
            # interp_factor = (time_idx - prev_save_time) / (next_save_time - prev_save_time)
            assert rec.replay_w[t] == pytest.approx((t - prev) / (nxt - prev), rel=1e-6)


def test_recorder_dtype_codes_and_buffer_shape():
    rec = fx.Recorder(modules=[fx.LinearReconstructEveryK(k=5), fx.DtypeConversion(dtype="float8_e4m3fnuz")]).init_tables(1311)
    assert rec._latent_array_size == len(range(0, 1311, 5)) and rec.dtype_code == 3
    assert fx.Recorder(modules=[fx.DtypeConversion(dtype="bfloat16")]).init_tables(10).dtype_code == 1
    assert fx.Recorder(modules=[]).init_tables(10)._latent_array_size == 10


# ---------------------------------------------------------------------------------------------
# TFSF per-component Yee time offsets (core/grid.py:745-880)
# ---------------------------------------------------------------------------------------------
def _expected_offsets(edges, center, k_hat, n_idx, dt):
    """-(r_q - center) . k_hat * n / (c0 dt) with r_q the Yee position of component q: E_q sits half a
    cell along q, H_q half a cell along the two other axes."""
    lo = [e[:-1] for e in edges]
    mid = [0.5 * (e[:-1] + e[1:]) for e in edges]
    out = {"E": [], "H": []}
    for fld in ("E", "H"):
        for q in range(3):
            ax = [(mid[a] if ((a == q) == (fld == "E")) else lo[a]) for a in range(3)]
            X, Y, Z = np.meshgrid(*ax, indexing="ij")
            proj = (X - center[0]) * k_hat[0] + (Y - center[1]) * k_hat[1] + (Z - center[2]) * k_hat[2]
            out[fld].append(-proj * n_idx / (C0 * dt))
    return np.stack(out["E"]), np.stack(out["H"])


def test_tfsf_time_offsets_uniform_tilted():
    cfg = fx.SimulationConfig(time=1e-14, grid=fx.UniformGrid(spacing=40e-9))
    face, sl = (7, 5, 1), ((2, 9), (3, 8), (4, 5))
    k = np.array([math.sin(0.3), 0.0, math.cos(0.3)])
    n_idx = 1.7
    center = [0.12e-6, 0.08e-6, 0.0]
    tE, tH = calculate_time_offset_yee(center, k.astype(F), np.full(face, n_idx, F), face, cfg, sl)
    edges = [np.arange(n + 1) * 40e-9 for n in face]
    eE, eH = _expected_offsets(edges, center, k, n_idx, cfg.time_step_duration)
    np.testing.assert_allclose(tE, eE, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(tH, eH, rtol=2e-5, atol=2e-5)
    # Ex and Hy sit half a cell further along x than Ey: one half-cell of travel along k_x
    half = 0.5 * 40e-9 * k[0] * n_idx / (C0 * cfg.time_step_duration)
    np.testing.assert_allclose(tE[1] - tE[0], half, rtol=1e-3)


def test_tfsf_time_offsets_nonuniform_use_physical_coordinates():
    rng = np.random.default_rng(0)
    ed = [np.concatenate([[0.0], np.cumsum(30e-9 * (1 + rng.random(n)))]) for n in (12, 10, 9)]
    cfg = fx.SimulationConfig(time=1e-14, grid=fx.RectilinearGrid(*ed))
    sl = ((4, 5), (1, 9), (2, 8))
    face = (1, 8, 6)
    k = np.array([1.0, 0.0, 0.0])
    center = [0.0, 0.1e-6, 0.07e-6]
    tE, tH = calculate_time_offset_yee(center, k.astype(F), np.full(face, 2.4, F), face, cfg, sl)
    local = [ed[a][sl[a][0]:sl[a][1] + 1] - ed[a][sl[a][0]] for a in range(3)]
    eE, eH = _expected_offsets(local, center, k, 2.4, cfg.time_step_duration)
    np.testing.assert_allclose(tE, eE, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(tH, eH, rtol=2e-5, atol=2e-5)
    # x-normal plane, k along x: only the components staggered along x are delayed (by half the local cell)
    w = ed[0][5] - ed[0][4]
    np.testing.assert_allclose(tE[0], -0.5 * w * 2.4 / (C0 * cfg.time_step_duration), rtol=1e-4)
    assert np.all(tE[1] == 0) and np.all(tH[0] == 0)


# ---------------------------------------------------------------------------------------------
# phasor detector tables (objects/detectors/phasor.py:100-235)
# ---------------------------------------------------------------------------------------------
def test_phasor_scale_window_and_table():
    from fdtdx_b200.detectors import phasor_table

    cfg = fx.SimulationConfig(time=4e-14, grid=fx.UniformGrid(spacing=50e-9))
    T, dt = cfg.time_steps_total, cfg.time_step_duration
    wcs = (fx.WaveCharacter(wavelength=1.55e-6), fx.WaveCharacter(wavelength=0.9e-6))
    det = fx.PhasorDetector(name="p", grid_slice_tuple=((1, 2), (0, 4), (0, 4)), wave_characters=wcs, switch=fx.OnOffSwitch(start_time=10 * dt))
    det.place_on_grid(cfg)
    n_on = sum(1 for t in range(T) if t * dt >= 10 * dt)
    assert det._window_at_time_step_arr.sum() == n_on
    assert det._static_scale() == pytest.approx(2.0 / n_on, rel=1e-12)  # continuous: 2 / sum(w)
    om = np.array([2 * math.pi * C0 / 1.55e-6, 2 * math.pi * C0 / 0.9e-6])
    np.testing.assert_allclose(det._angular_frequencies, om, rtol=1e-6)
    tab = phasor_table(det, T, dt)  # (T, nf, 2): exp(+i w t dt)
    for t in (0, 1, T // 2, T - 1):
        np.testing.assert_allclose(tab[t, :, 0], np.cos(om * t * dt), atol=3e-4)
        np.testing.assert_allclose(tab[t, :, 1], np.sin(om * t * dt), atol=3e-4)
    # pulse scaling: scale = stride, on-list thinned to every stride-th ACTIVE step (phasor.py:104-116)
    det = fx.ModeOverlapDetector(name="q", grid_slice_tuple=((1, 2), (0, 4), (0, 4)), wave_characters=wcs[:1], scaling_mode="pulse", dft_subsample=3,
                                 switch=fx.OnOffSwitch(start_time=10 * dt))
    det.place_on_grid(cfg)
    active = [t for t in range(T) if t * dt >= 10 * dt]
    assert [t for t in range(T) if det._is_on_at_time_step_arr[t]] == active[::3]
    assert det._static_scale() == 3
    # "auto": floor(1 / (8 f_max dt))
    det = fx.PhasorDetector(name="r", grid_slice_tuple=((1, 2), (0, 4), (0, 4)), wave_characters=wcs, dft_subsample="auto")
    det.place_on_grid(cfg)
    assert det._dft_stride == max(1, math.floor(1.0 / (8.0 * (C0 / 0.9e-6) * dt)))


# ---------------------------------------------------------------------------------------------
# metric scales (core/physics/curl.py:10-39), time step (core/grid.py:495-511)
# ---------------------------------------------------------------------------------------------
def test_metric_scales_and_cfl_time_step():
    rng = np.random.default_rng(1)
    ed = [np.concatenate([[0.0], np.cumsum(25e-9 * (1 + rng.random(n)))]) for n in (9, 8, 7)]
    cfg = fx.SimulationConfig(time=1e-14, grid=fx.RectilinearGrid(*ed), courant_factor=0.9)
    w = [np.diff(e) for e in ed]
    dt = 0.9 / (C0 * math.sqrt(sum(1.0 / wi.min() ** 2 for wi in w)))
    assert cfg.time_step_duration == pytest.approx(dt, rel=1e-12)
    ref = C0 * dt / (0.9 / math.sqrt(3))
    for a in range(3):
        sB, sF = metric_scales(cfg, a)
        prev = np.concatenate([w[a][:1], w[a][:-1]])
        np.testing.assert_allclose(sF, ref / w[a], rtol=1e-6)
        np.testing.assert_allclose(sB, ref / (0.5 * (w[a] + prev)), rtol=1e-6)
    cfg = fx.SimulationConfig(time=1e-14, grid=fx.UniformGrid(spacing=20e-9), courant_factor=0.99)
    assert cfg.time_step_duration == pytest.approx(0.99 / math.sqrt(3) * 20e-9 / C0, rel=1e-12)
    assert fx.SimulationConfig(time=350e-15, grid=fx.UniformGrid(spacing=20e-9), courant_factor=0.99).time_steps_total == 9179


# ---------------------------------------------------------------------------------------------
# ADE coefficients and the dispersive source set-up (dispersion.py:789-885, 988-1064, 1180-1307)
# ---------------------------------------------------------------------------------------------
def test_lorentz_coefficients_and_recovered_susceptibility():
    w0, g, de, dt = 3.93282466e15, 1e8, 3.68799143, 3.81e-17
    c1, c2, c3, c4 = dispersion.pole_coefficients((dispersion.LorentzPole(w0, g, de),), dt)
    D = 1 + g * dt / 2
    assert c1[0] == pytest.approx((2 - (w0 * dt) ** 2) / D, rel=1e-14)
    assert c2[0] == pytest.approx(-(1 - g * dt / 2) / D, rel=1e-14)
    assert c3[0] == pytest.approx(de * w0**2 * dt**2 / D, rel=1e-14)
    assert c4[0] == 0.0
    # chi recovered from the coefficients is the analytic Lorentzian de w0^2 / (w0^2 - w^2 - i g w)
    om = 2 * math.pi * C0 / 1.55e-6
    chi = dispersion.susceptibility(c1[:, None], c2[:, None], c3[:, None], om, dt)
    exact = de * w0**2 / (w0**2 - om**2 - 1j * g * om)
    assert complex(chi[0]) == pytest.approx(exact, rel=1e-9)
    # Drude: omega_0 = 0, a = wp^2
    c1, c2, c3, _ = dispersion.pole_coefficients((dispersion.DrudePole(2e15, 1e13),), dt)
    assert c1[0] == pytest.approx(2 / (1 + 1e13 * dt / 2), rel=1e-14) and c3[0] == pytest.approx(4e30 * dt**2 / (1 + 1e13 * dt / 2), rel=1e-14)
    with pytest.raises(ValueError):
        dispersion.pole_coefficients((dispersion.LorentzPole(3e17, 0.0, 1.0),), 1e-17)


def test_dispersive_H_filter_is_identity_without_dispersion_and_matches_direct_convolution():
    dt, T = 4e-17, 600
    t = np.arange(T) * dt
    raw = np.exp(-(((t - 6e-15) / 2e-15) ** 2)) * np.cos(2 * math.pi * 1.9e14 * t)
    shape = (1, 1, 2, 2, 1)
    z = np.zeros(shape)
    inv_eps = np.full((1, 2, 2, 1), 1 / 4.0)
    assert np.array_equal(dispersion.dispersive_H_filter(raw, dt, z, z, z, inv_eps, 1.2e15), raw)
    c = dispersion.pole_coefficients((dispersion.LorentzPole(3.9e15, 1e12, 2.0),), dt)
    c1, c2, c3 = (np.full(shape, v[0]) for v in c[:3])
    out = dispersion.dispersive_H_filter(raw, dt, c1, c2, c3, inv_eps, 1.2e15)
    # direct evaluation: zero-pad to M, multiply the spectrum by sqrt(eps(w)/eps(wc)), DC gain 1
    M = 2048
    om = 2 * math.pi * np.fft.rfftfreq(M, d=dt)
    eps = lambda w: 4.0 + 2.0 * 3.9e15**2 / (3.9e15**2 - w**2 - 1j * 1e12 * w)
    # the discrete recurrence realises the pole with its own (slightly warped) frequency response; at
    # omega*dt << 1 the analytic Lorentzian agrees to ~1e-3, which bounds this comparison
    G = np.sqrt(eps(om) / eps(1.2e15))
    G[0] = 1.0
    G[-1] = G[-1].real
    pad = np.zeros(M)
    pad[:T] = raw
    exp = np.fft.irfft(np.fft.rfft(pad) * G, n=M)[:T]
    assert np.linalg.norm(out - exp) / np.linalg.norm(exp) < 5e-3
    assert np.linalg.norm(out - raw) / np.linalg.norm(raw) > 1e-3  # the filter does something


def test_effective_inv_permittivity_at_the_carrier():
    dt = 3.8e-17
    pole = dispersion.LorentzPole(3.93282466e15, 1e8, 3.68799143)
    co = dispersion.coefficient_arrays((pole,), dt, np.ones((2, 2, 2), bool))
    inv = np.full((1, 2, 2, 2), 1 / 7.98737492, F)
    om = 2 * math.pi * C0 / 1.55e-6
    eff = dispersion.effective_inv_permittivity(inv, co["c1"], co["c2"], co["c3"], om, dt)
    exact = 7.98737492 + (3.68799143 * 3.93282466e15**2 / (3.93282466e15**2 - om**2 - 1j * 1e8 * om)).real
    np.testing.assert_allclose(1.0 / eff, exact, rtol=2e-5)  # float32 coefficients
    assert 12.0 < exact < 12.2  # silicon at 1.55 um


# ---------------------------------------------------------------------------------------------
# mode solver (core/physics/modes.py:101-385 + tidy3d's published FD eigen-problem)
# ---------------------------------------------------------------------------------------------
def _slab_neff(n_core, n_clad, wl, thickness, tm=False):
    from scipy.optimize import brentq

    k0 = 2 * math.pi / wl
    r = (n_core / n_clad) ** 2 if tm else 1.0

    def f(ne):
        kx, gm = k0 * math.sqrt(n_core**2 - ne**2), k0 * math.sqrt(ne**2 - n_clad**2)
        return math.tan(kx * thickness / 2) - r * gm / kx

    # fundamental even mode: kx d / 2 < pi / 2
    ne_min = math.sqrt(max(n_core**2 - (math.pi / (k0 * thickness)) ** 2, n_clad**2)) + 1e-9
    return brentq(f, ne_min, n_core - 1e-9)


def test_mode_solver_slab_effective_indices_and_normalisation():
    from fdtdx_b200 import modes

    wl, dx = 1.55e-6, 10e-9
    ny, nz = 3, 400
    eps = np.full((ny, nz), 2.25)
    zc = (np.arange(nz) + 0.5) * dx
    eps[:, np.abs(zc - nz * dx / 2) <= 110e-9] = 12.25
    inv = (1 / eps)[None, None].astype(F)  # (1, 1, ny, nz): propagation along x
    f = C0 / wl
    E, H, n_te = modes.compute_mode(f, inv, 1.0, resolution=dx, mode_index=0, filter_pol="te")
    assert n_te.real == pytest.approx(_slab_neff(3.5, 1.5, wl, 220e-9), rel=2e-3)
    assert abs(n_te.imag) < 1e-9
    # TE of this slab: E along y (first transverse axis), H in (x, z)
    en = [float(np.sum(np.abs(E[c]) ** 2)) for c in range(3)]
    assert en[1] > 1e6 * max(en[0], en[2])
    # unit Poynting flux along +x: sum 0.5 Re(E x H*)_x = 1 (metrics.py:163-219)
    S = 0.5 * np.real(np.cross(np.conj(E), H, axisa=0, axisb=0, axisc=0)[0]).sum()
    assert S == pytest.approx(1.0, rel=1e-9)
    _, _, n_tm = modes.compute_mode(f, inv, 1.0, resolution=dx, mode_index=0, filter_pol="tm")
    assert n_tm.real == pytest.approx(_slab_neff(3.5, 1.5, wl, 220e-9, tm=True), rel=2e-3)
    # backward mode: same index, reversed power flow
    Eb, Hb, n_b = modes.compute_mode(f, inv, 1.0, resolution=dx, direction="-", mode_index=0, filter_pol="te")
    Sb = 0.5 * np.real(np.cross(np.conj(Eb), Hb, axisa=0, axisb=0, axisc=0)[0]).sum()
    assert n_b.real == pytest.approx(n_te.real, rel=1e-9) and Sb == pytest.approx(-1.0, rel=1e-9)


def test_mode_solver_nonuniform_grid_and_other_axes():
    from fdtdx_b200 import modes

    wl = 1.55e-6
    f = C0 / wl
    # the same slab on a stretched z grid (fine in the core), propagation along y this time
    zf = np.concatenate([np.linspace(0, 1.2e-6, 31)[:-1], np.linspace(1.2e-6, 1.8e-6, 61)[:-1], np.linspace(1.8e-6, 3.0e-6, 31)])
    zc = 0.5 * (zf[:-1] + zf[1:])
    nx = 3
    eps = np.full((nx, len(zc)), 2.25)
    eps[:, np.abs(zc - 1.5e-6) <= 110e-9] = 12.25
    inv = (1 / eps)[None, :, None, :].astype(F)  # (1, nx, 1, nz)
    xe = np.arange(nx + 1) * 20e-9
    E, H, n = modes.compute_mode(f, inv, 1.0, mode_index=0, filter_pol="te", transverse_coords=[xe, zf])
    assert n.real == pytest.approx(_slab_neff(3.5, 1.5, wl, 220e-9), rel=5e-3)
    en = [float(np.sum(np.abs(E[c]) ** 2)) for c in range(3)]
    assert en[0] > 1e6 * max(en[1], en[2])  # E along x, the first transverse axis
    area = np.diff(xe)[:, None] * np.diff(zf)[None, :]
    S = 0.5 * np.real(np.cross(np.conj(E), H, axisa=0, axisb=0, axisc=0)[1])[:, 0, :]
    assert (S * area / area.mean()).sum() == pytest.approx(1.0, rel=1e-9)
    assert S.sum() > 0  # +y propagation carries power along +y in the right-handed physical frame
