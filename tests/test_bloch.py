"""Bloch boundaries with a non-zero wave vector: complex fields.

Reference: ``objects/boundaries/bloch.py:31-155`` (phase, ghost-plane correction),
``fdtd/initialization.py:581-596`` (complex promotion), ``tests/unit/objects/boundaries/test_bloch.py:120-210``
(phase known answers, re-typed here), ``tests/simulation/fdtd/test_time_reversal.py:497-553`` (complex
time reversal with interface recording).  CPU tests pin the oracle; ``-m gpu`` tests compare the two-real-
systems CUDA path (``fdtdx_b200/bloch.py``) with the complex64 oracle through the public drivers."""

import numpy as np
import pytest

import fdtdx_b200 as fx
from oracle import yee
from scenes import make_config, rel_l2

F = np.float32
BXY = {"min_x": "bloch", "max_x": "bloch", "min_y": "bloch", "max_y": "bloch", "min_z": "pml", "max_z": "pml"}
BZ = {"min_x": "periodic", "max_x": "periodic", "min_y": "pml", "max_y": "pml", "min_z": "bloch", "max_z": "bloch"}


def build(shape, types, k, time=6e-15, seed=0, recorder=None, source=False, detectors=(), nonuniform=False, thickness=3):
    cfg = make_config(shape, time=time, recorder=recorder, nonuniform=nonuniform, seed=seed)
    nx, ny, nz = shape
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=((0, nx), (0, ny), (0, nz)))
    bl = fx.boundary_objects_from_config(shape, cfg, types, thickness=thickness, bloch_vector=k)
    rng = np.random.default_rng(seed + 1)
    inv_eps = (1.0 / (1.0 + 3.0 * rng.random((1, *shape)))).astype(F)
    objs = [vol, *bl]
    wc = fx.WaveCharacter(wavelength=0.8e-6)
    pml_axis = next(b.axis for b in bl if isinstance(b, fx.PerfectlyMatchedLayer))
    n_ax = shape[pml_axis]
    if source:
        sl = [(0, nx), (0, ny), (0, nz)]
        sl[pml_axis] = (thickness + 1, thickness + 2)
        pol = (1.0, 0.0, 0.0) if pml_axis != 0 else (0.0, 1.0, 0.0)
        objs.append(fx.make_plane_source("source", tuple(sl), cfg, inv_eps, 1.0, direction="+", wave_character=wc, fixed_E_polarization_vector=pol))
    plane = [(0, nx), (0, ny), (0, nz)]
    plane[pml_axis] = (n_ax - thickness - 3, n_ax - thickness - 2)
    full = ((0, nx), (0, ny), (0, nz))
    mk = {
        "energy": lambda: fx.EnergyDetector(name="energy", grid_slice_tuple=full, switch=fx.OnOffSwitch(interval=2)),
        "energy_slices": lambda: fx.EnergyDetector(name="energy_slices", grid_slice_tuple=full, as_slices=True),
        "poynting": lambda: fx.PoyntingFluxDetector(name="poynting", grid_slice_tuple=tuple(plane), direction="+", fixed_propagation_axis=pml_axis),
        "phasor": lambda: fx.PhasorDetector(name="phasor", grid_slice_tuple=tuple(plane), wave_characters=(wc,)),
        "field": lambda: fx.FieldDetector(name="field", grid_slice_tuple=tuple(plane), components=("Ex", "Hy")),
    }
    for d in detectors:
        objs.append(mk[d]())
    objects, arrays, _, cfg, _ = fx.place_objects(objs, cfg, inv_permittivities=inv_eps)
    return objects, arrays, cfg


def seed_complex(arrays, seed=0, amp=1e-3):
    rng = np.random.default_rng(seed)
    cn = lambda shp: (amp * (rng.standard_normal(shp) + 1j * rng.standard_normal(shp))).astype(np.complex64)
    arrays.fields.E[...] = cn(arrays.fields.E.shape)
    arrays.fields.H[...] = cn(arrays.fields.H.shape)
    for d in (arrays.fields.psi_E, arrays.fields.psi_H):
        for k, (a, b) in d.items():
            a[...] = 0.1 * cn(a.shape)
            b[...] = 0.1 * cn(b.shape)
    return arrays


# ------------------------------------------------------------------------------------------ CPU: the oracle
def test_bloch_phase_known_answers():
    """test_bloch.py:120-210: zero vector -> unit phase; |phase| = 1; phase = exp(i k_axis L); only the
    component along the boundary's own axis counts."""
    shape = (8, 6, 12)
    cfg = make_config(shape)
    b0 = fx.BlochBoundary(name="b", grid_slice_tuple=((0, 1), (0, 6), (0, 12)), axis=0, direction="-", bloch_vector=(0.0, 0.0, 0.0))
    assert not b0.needs_complex_fields and b0.get_bloch_phase(shape, cfg) == np.complex64(1.0)
    k = (2.0e6, 3.0e6, 0.0)
    for axis in range(3):
        b = fx.BlochBoundary(name="b", grid_slice_tuple=((0, 8), (0, 6), (0, 12)), axis=axis, direction="+", bloch_vector=k)
        ph = b.get_bloch_phase(shape, cfg)
        assert ph.dtype == np.complex64 and abs(abs(ph) - 1.0) < 1e-6
        L = shape[axis] * cfg.uniform_spacing()
        assert abs(ph - np.exp(1j * k[axis] * L)) < 1e-6
        assert b.needs_complex_fields == (k[axis] != 0.0)
    cfg_nu = make_config(shape, nonuniform=True)
    b = fx.BlochBoundary(name="b", grid_slice_tuple=((0, 8), (0, 6), (0, 12)), axis=1, direction="+", bloch_vector=k)
    e = cfg_nu.resolved_grid.edges(1)
    assert abs(b.get_bloch_phase(shape, cfg_nu) - np.exp(1j * k[1] * (e[6] - e[0]))) < 1e-5  # physical extent of the axis


def test_pad_correction_multiplies_the_wrapped_ghost_planes():
    """bloch.py:61-96: '-' face ghost = F[N-1] * conj(phase), '+' face ghost = F[0] * phase; corners get both."""
    objects, arrays, cfg = build((6, 5, 10), BXY, (1.3e6, 0.7e6, 0.0))
    assert arrays.fields.E.dtype == np.complex64 and arrays.fields.psi_E[objects.pml_objects[0].name][0].dtype == np.complex64
    seed_complex(arrays)
    E = arrays.fields.E
    pad = yee.pad_fields_for_boundaries(E, objects, cfg)
    px = next(b for b in objects.boundary_objects if getattr(b, "needs_complex_fields", False) and b.axis == 0).get_bloch_phase((6, 5, 10), cfg)
    py = next(b for b in objects.boundary_objects if getattr(b, "needs_complex_fields", False) and b.axis == 1).get_bloch_phase((6, 5, 10), cfg)
    np.testing.assert_allclose(pad[:, 0, 1:-1, 1:-1], E[:, -1] * np.conj(px), rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(pad[:, -1, 1:-1, 1:-1], E[:, 0] * px, rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(pad[:, 1:-1, 0, 1:-1], E[:, :, -1] * np.conj(py), rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(pad[:, 0, -1, 1:-1], E[:, -1, 0] * np.conj(px) * py, rtol=1e-5, atol=1e-10)
    assert np.all(pad[:, :, :, 0] == 0) and np.all(pad[:, :, :, -1] == 0)  # PML axis: zero halo


def test_antiperiodic_supercell_equals_bloch_half_cell():
    """Known answer built from the pinned k = 0 path only: a periodic domain of 2N planes seeded with
    F[N+i] = -F[i] (materials of period N) stays antiperiodic, so its first N planes evolve exactly like the
    N-plane Bloch cell with k L = pi (phase -1)."""
    N, ny, nz = 6, 4, 10
    PER = {"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}
    BX = dict(PER, min_x="bloch", max_x="bloch")
    o2, a2, c2 = build((2 * N, ny, nz), PER, (0.0, 0.0, 0.0))
    L = N * c2.uniform_spacing()
    o1, a1, c1 = build((N, ny, nz), BX, (np.pi / L, 0.0, 0.0))
    rng = np.random.default_rng(3)
    inv_eps = (1.0 / (1.0 + 3.0 * rng.random((1, N, ny, nz)))).astype(F)
    a1 = a1.aset("inv_permittivities", inv_eps)
    a2 = a2.aset("inv_permittivities", np.concatenate([inv_eps, inv_eps], axis=1))
    for name in ("E", "H"):
        v = (1e-3 * rng.standard_normal((3, N, ny, nz))).astype(F)
        getattr(a1.fields, name)[...] = v.astype(np.complex64)
        getattr(a2.fields, name)[...] = np.concatenate([v, -v], axis=1)
    s1, s2 = (0, a1), (0, a2)
    for _ in range(12):
        s1 = yee.forward(s1, c1, o1, record_detectors=False)
        s2 = yee.forward(s2, c2, o2, record_detectors=False)
    for name in ("E", "H"):
        got, ref = getattr(s1[1].fields, name), getattr(s2[1].fields, name)
        assert rel_l2(got.real, ref[:, :N]) <= 2e-6, name                       # float32 sin(pi) leaks ~1e-7 into Im
        assert np.abs(got.imag).max() <= 1e-5 * np.abs(got.real).max(), name
        np.testing.assert_allclose(ref[:, N:], -ref[:, :N], rtol=0, atol=1e-9)    # the supercell stayed antiperiodic


def test_complex_time_reversal_oracle():
    """test_time_reversal.py:497-529: forward steps with interface recording, then backward steps, recover the
    seeded complex fields (atol 1e-5 on fields of size ~1e-3 there; relative 1e-4 here)."""
    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build((8, 6, 14), BXY, (1e6, 1e6, 0.0), recorder=rec, time=4e-15)
    rng = np.random.default_rng(0)
    inner = (slice(None), slice(None), slice(None), slice(4, -4))
    for name in ("E", "H"):  # seeded away from the PML, like _seed_fields
        v = getattr(arrays.fields, name)
        v[inner] = (1e-3 * (rng.standard_normal(v[inner].shape) + 1j * rng.standard_normal(v[inner].shape))).astype(np.complex64)
    E0, H0 = arrays.fields.E.copy(), arrays.fields.H.copy()
    st = (0, arrays)
    n = 6
    for _ in range(n):
        st = yee.forward(st, cfg, objects, record_detectors=False, record_boundaries=True)
    assert all(np.iscomplexobj(v) for v in st[1].recording_state.data.values())
    for _ in range(n):
        st = yee.backward(st, cfg, objects, record_detectors=False, reset_fields=True)
    assert rel_l2(st[1].fields.E[inner], E0[inner]) <= 1e-4
    assert rel_l2(st[1].fields.H[inner], H0[inner]) <= 1e-4


# ------------------------------------------------------------------------------------------ GPU: CUDA vs oracle
def _np(t):
    return t.detach().cpu().numpy()


CASES = {
    "bloch_xy_pml_z": dict(shape=(10, 8, 16), types=BXY, k=(1e6, 1e6, 0.0)),
    "bloch_x_only": dict(shape=(12, 6, 16), types=dict(BXY, min_y="periodic", max_y="periodic"), k=(2.1e6, 0.0, 0.0)),
    "bloch_z_ragged": dict(shape=(6, 12, 9), types=BZ, k=(0.0, 0.0, 1.7e6)),
    "bloch_xy_nonuniform": dict(shape=(8, 8, 20), types=BXY, k=(0.8e6, -1.2e6, 0.0), nonuniform=True),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_complex_forward_matches_oracle(case):
    """Seeded complex state + a plane source, all detector kinds, 10 steps: fields rel-L2 <= 1e-5, detectors <= 1e-4."""
    kw = dict(CASES[case])
    objects, arrays, cfg = build(kw.pop("shape"), kw.pop("types"), kw.pop("k"), source=True, time=5e-15,
                                 detectors=("energy", "energy_slices", "poynting", "phasor", "field"), **kw)
    seed_complex(arrays, seed=4)
    T = min(10, cfg.time_steps_total)
    st_o = (0, arrays.map_arrays(lambda a: a.copy() if isinstance(a, np.ndarray) else a))
    for _ in range(T):
        st_o = yee.forward(st_o, cfg, objects, record_detectors=True)
    dev = arrays.to_torch("cuda")
    t, out = fx.custom_fdtd_forward(dev, objects, cfg, reset_container=False, record_detectors=True, start_time=0, end_time=T)
    assert t == T and out.fields.E.is_complex()
    for name in ("E", "H"):
        e = rel_l2(_np(getattr(out.fields, name)), getattr(st_o[1].fields, name))
        assert e <= 1e-5, (case, name, e)
    for pml in objects.pml_objects:
        for w in range(2):
            assert rel_l2(_np(out.fields.psi_E[pml.name][w]), st_o[1].fields.psi_E[pml.name][w]) <= 1e-5
            assert rel_l2(_np(out.fields.psi_H[pml.name][w]), st_o[1].fields.psi_H[pml.name][w]) <= 1e-5
    for d, st in st_o[1].detector_states.items():
        for key, ref in st.items():
            got = _np(out.detector_states[d][key])
            e = rel_l2(got, ref)
            assert e <= 1e-4, (case, d, key, e)


@pytest.mark.gpu
def test_complex_time_reversal_cuda():
    """forward(record_boundaries) x n then backward x n through the public single-step drivers: the recordings
    are complex, CUDA == oracle after the round trip, and the interior is reconstructed."""
    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build((8, 6, 14), BXY, (1e6, 1e6, 0.0), recorder=rec, time=4e-15, detectors=("energy",))
    rng = np.random.default_rng(0)
    inner = (slice(None), slice(None), slice(None), slice(4, -4))
    for name in ("E", "H"):
        v = getattr(arrays.fields, name)
        v[inner] = (1e-3 * (rng.standard_normal(v[inner].shape) + 1j * rng.standard_normal(v[inner].shape))).astype(np.complex64)
    E0 = arrays.fields.E.copy()
    n = 6
    st_o = (0, arrays.map_arrays(lambda a: a.copy() if isinstance(a, np.ndarray) else a))
    st_g = (0, arrays.to_torch("cuda"))
    for _ in range(n):
        st_o = yee.forward(st_o, cfg, objects, record_detectors=True, record_boundaries=True)
        st_g = fx.forward(st_g, cfg, objects, record_detectors=True, record_boundaries=True)
    for k, v in st_g[1].recording_state.data.items():
        assert v.is_complex() and rel_l2(_np(v), st_o[1].recording_state.data[k]) <= 1e-5, k
    assert rel_l2(_np(st_g[1].fields.E), st_o[1].fields.E) <= 1e-5
    st_o = (n, st_o[1])
    for _ in range(n):
        st_o = yee.backward(st_o, cfg, objects, record_detectors=False, reset_fields=True)
    st_g = fx.full_backward(st_g, objects, cfg, record_detectors=False, reset_fields=True)
    assert st_g[0] == 0
    for name in ("E", "H"):
        assert rel_l2(_np(getattr(st_g[1].fields, name)), getattr(st_o[1].fields, name)) <= 1e-5, name
    assert rel_l2(_np(st_g[1].fields.E)[inner], E0[inner]) <= 1e-4
