"""CUDA path vs the CPU oracle on identical seeded inputs (the parity gate).

Tolerances (BASELINE.json north_star): fields after N steps rel-L2 <= 1e-5; detector outputs
<= 1e-4.  The kernels are compiled without FMA contraction and follow the oracle's float32 op
order, so field parity is in practice ~1e-7; the stated bounds are what is asserted.
"""

import numpy as np
import pytest

import fdtdx_b200 as fx
from oracle import yee
from scenes import build_scene, rel_l2, seed_fields

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-5
DET_TOL = 1e-4


def _to_np(x):
    import torch

    if isinstance(x, torch.Tensor):
        return x.detach().to(torch.float32).cpu().numpy() if not x.is_complex() else x.detach().cpu().numpy()
    return np.asarray(x)


def run_both(objects, arrays, cfg, steps, seed=True, record_boundaries=False, simulate_boundaries=True, record_detectors=True):
    if seed:
        seed_fields(arrays, seed=3)
    dev = arrays.to_torch("cuda")
    st_o = (0, arrays)
    for _ in range(steps):
        st_o = yee.forward(st_o, cfg, objects, None, record_detectors, record_boundaries, simulate_boundaries)
    st_g = (0, dev)
    from fdtdx_b200.fdtd import get_plan

    plan = get_plan(dev, objects, cfg)
    plan.run_forward(0, steps, record_detectors, record_boundaries, simulate_boundaries)
    out = plan.finish(dev)
    return st_o[1], out


def assert_fields_close(a_o, a_g, tol=FIELD_TOL):
    assert rel_l2(_to_np(a_g.fields.E), a_o.fields.E) <= tol, "E mismatch"
    assert rel_l2(_to_np(a_g.fields.H), a_o.fields.H) <= tol, "H mismatch"
    for name in a_o.fields.psi_E:
        for w in range(2):
            assert rel_l2(_to_np(a_g.fields.psi_E[name][w]), a_o.fields.psi_E[name][w]) <= tol, f"psi_E {name}"
            assert rel_l2(_to_np(a_g.fields.psi_H[name][w]), a_o.fields.psi_H[name][w]) <= tol, f"psi_H {name}"
    if a_o.fields.dispersive_P_curr is not None:
        assert rel_l2(_to_np(a_g.fields.dispersive_P_curr), a_o.fields.dispersive_P_curr) <= tol, "P_curr"
        assert rel_l2(_to_np(a_g.fields.dispersive_P_prev), a_o.fields.dispersive_P_prev) <= tol, "P_prev"


def assert_detectors_close(a_o, a_g, tol=DET_TOL):
    for name, st in a_o.detector_states.items():
        for key, ref in st.items():
            got = _to_np(a_g.detector_states[name][key])
            assert got.shape == ref.shape, (name, key, got.shape, ref.shape)
            assert rel_l2(got, ref) <= tol, f"detector {name}/{key}: {rel_l2(got, ref)}"


CASES = {
    # name: (kwargs, steps)
    "vacuum_periodic": (dict(boundaries="periodic"), 12),
    "pml_all": (dict(boundaries="pml", thickness=3), 12),
    "pml_kappa": (dict(boundaries="pml", thickness=3, kappa=True), 8),
    "pec_pmc": (dict(boundaries={"min_x": "pec", "max_x": "pmc", "min_y": "pmc", "max_y": "pec", "min_z": "pec", "max_z": "pml"}, thickness=3), 10),
    "mixed_periodic_pml": (dict(boundaries={"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}, thickness=4), 10),
    "diag_eps": (dict(eps_tier=3), 8),
    "sigma": (dict(sigma_E=True, sigma_H=True), 8),
    "diag_sigma_mu": (dict(eps_tier=3, sigma_E=True, sigma_H=True, mu_tier=3), 8),
    "mu_iso": (dict(mu_tier=1), 6),
    "nonuniform": (dict(nonuniform=True), 10),
    "nonuniform_all": (dict(nonuniform=True, eps_tier=3, sigma_E=True, mu_tier=1, sigma_H=True), 8),
    "scalar_path_odd_nz": (dict(shape=(9, 7, 13)), 8),
    # ragged rows (Nz % 4 != 0): four cells per thread, predicated 32-bit accesses, partial last lane
    "ragged_nz65_metric": (dict(shape=(10, 11, 65), thickness=4, nonuniform=True, eps_tier=3), 6),
    "ragged_nz3": (dict(shape=(6, 5, 3), thickness=1, boundaries={"min_x": "pml", "max_x": "pml", "min_y": "pml", "max_y": "pml", "min_z": "pec", "max_z": "pmc"}), 6),
    "ragged_periodic_nz13": (dict(shape=(5, 6, 13), boundaries="periodic"), 8),
    "ragged_periodic_z_nz134": (dict(shape=(4, 5, 134), thickness=2, boundaries={"min_x": "pml", "max_x": "pml", "min_y": "pml", "max_y": "pml", "min_z": "periodic", "max_z": "periodic"}), 6),
    "ragged_all_tiers_nz18": (dict(shape=(6, 7, 18), nonuniform=True, eps_tier=3, sigma_E=True, mu_tier=3, sigma_H=True, kappa=True), 6),
    "ragged_ade_nz10": (dict(shape=(6, 6, 10), poles=2, c4=True, eps_tier=3, coeff_tier=3), 6),
    "wide_z": (dict(shape=(6, 9, 260), thickness=2), 6),
    "multi_chunk": (dict(shape=(70, 9, 8), thickness=2), 6),
    "ade_1pole": (dict(poles=1), 8),
    "ade_2pole_c4_sigma": (dict(poles=2, c4=True, sigma_E=True, eps_tier=3, coeff_tier=3), 8),
    "ade_sigma_noc4": (dict(poles=1, sigma_E=True), 6),
}


@pytest.mark.parametrize("name", list(CASES))
def test_seeded_step_parity(name):
    kw, steps = CASES[name]
    objects, arrays, cfg = build_scene(**kw)
    a_o, a_g = run_both(objects, arrays, cfg, steps)
    assert_fields_close(a_o, a_g)


@pytest.mark.parametrize("name", [n for n in CASES if n.startswith("ragged") or n == "scalar_path_odd_nz"])
def test_ragged_kernels_without_z_padding(name, monkeypatch):
    """Forward-only runs on ragged rows normally go through z-padded shadows (plan.py) and the staged
    kernels; this forces the interleaved ragged marching kernels on the same scenes."""
    monkeypatch.setenv("FDTDX_B200_PAD_Z", "0")
    kw, steps = CASES[name]
    objects, arrays, cfg = build_scene(**kw)
    a_o, a_g = run_both(objects, arrays, cfg, steps)
    assert_fields_close(a_o, a_g)


def test_z_padding_leaves_no_trace():
    """Padded shadows: results on a ragged grid are bit-identical with and without the padding path
    (same arithmetic; the padded cells are the zero halo), detectors and sources included."""
    import os

    outs = []
    for pad in ("1", "0"):
        os.environ["FDTDX_B200_PAD_Z"] = pad
        try:
            objects, arrays, cfg = build_scene(shape=(14, 11, 21), thickness=3, source="plane_x", nonuniform=True, eps_tier=3,
                                               detectors=("energy_slices", "phasor", "poynting"), time=6e-15)
            steps = min(cfg.time_steps_total, 40)
            _, a_g = run_both(objects, arrays, cfg, steps, seed=False)
            outs.append(a_g)
        finally:
            os.environ.pop("FDTDX_B200_PAD_Z", None)
    assert np.array_equal(_to_np(outs[0].fields.E), _to_np(outs[1].fields.E))
    assert np.array_equal(_to_np(outs[0].fields.H), _to_np(outs[1].fields.H))
    for name in outs[0].fields.psi_E:
        for w in range(2):
            assert np.array_equal(_to_np(outs[0].fields.psi_E[name][w]), _to_np(outs[1].fields.psi_E[name][w])), name
    for name, st in outs[0].detector_states.items():
        for key, v in st.items():
            assert np.array_equal(_to_np(v), _to_np(outs[1].detector_states[name][key])), (name, key)


TENSOR_CASES = {
    "eps9_pml": (dict(eps_tier=9), 8),
    "eps9_periodic": (dict(eps_tier=9, boundaries="periodic"), 8),
    "eps9_sigma9": (dict(eps_tier=9, sigma_E=9), 6),
    "eps9_mu9": (dict(eps_tier=9, mu_tier=9), 6),
    "mu9_only": (dict(mu_tier=9, eps_tier=3), 6),
    "eps9_nonuniform": (dict(eps_tier=9, nonuniform=True), 6),
    "eps9_pec": (dict(eps_tier=9, boundaries={"min_x": "pec", "max_x": "pmc", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}), 6),
    "eps9_ade": (dict(eps_tier=9, poles=2, coeff_tier=3), 6),
    "eps9_odd_steps": (dict(eps_tier=9, mu_tier=9, shape=(7, 9, 11)), 5),
}


@pytest.mark.parametrize("name", list(TENSOR_CASES))
def test_full_tensor_step_parity(name):
    """Full-tensor tier (update.py:356-492, 752-822): two-phase curl-scratch + 3x3 apply kernels."""
    kw, steps = TENSOR_CASES[name]
    objects, arrays, cfg = build_scene(**kw)
    a_o, a_g = run_both(objects, arrays, cfg, steps)
    assert_fields_close(a_o, a_g)


@pytest.mark.parametrize("src", ["plane_z", "plane_x", "dipole"])
def test_full_tensor_sources_and_detectors(src):
    """TFSF injection into all three rows under full anisotropy (tfsf.py:285-295, 384-391)."""
    shape = (16, 10, 12) if src == "plane_x" else (12, 10, 16)
    objects, arrays, cfg = build_scene(shape=shape, eps_tier=9, mu_tier=9 if src == "plane_z" else 0, source=src,
                                       detectors=("field", "phasor", "poynting", "energy", "energy_reduce"), time=6e-15)
    steps = min(cfg.time_steps_total, 40)
    a_o, a_g = run_both(objects, arrays, cfg, steps, seed=False)
    assert np.abs(a_o.fields.E).max() > 0
    assert_fields_close(a_o, a_g)
    assert_detectors_close(a_o, a_g)


def test_full_tensor_reverse():
    """C3b-style: forward with recording then reverse steps through the tensor reverse update
    (update.py:610-681, 932-1002; A = M2^-1 M1)."""
    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build_scene(shape=(14, 12, 16), thickness=4, eps_tier=9, sigma_E=9, source="plane_z", recorder=rec, time=4e-15,
                                       boundaries={"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"})
    T = cfg.time_steps_total
    st_o = yee.checkpointed_fdtd(arrays, objects, cfg)
    st_g = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg)
    assert_fields_close(st_o[1], st_g[1])
    st_o = yee.full_backward(st_o, objects, cfg, record_detectors=False, reset_fields=True, start_time_step=T - 7)
    st_g = fx.full_backward(st_g, objects, cfg, record_detectors=False, reset_fields=True, start_time_step=T - 7)
    assert_fields_close(st_o[1], st_g[1], tol=DET_TOL)


@pytest.mark.parametrize("periodic", [True, False])
def test_simulate_boundaries_false(periodic):
    objects, arrays, cfg = build_scene(boundaries="periodic" if periodic else "pml")
    a_o, a_g = run_both(objects, arrays, cfg, 5, simulate_boundaries=False)
    assert_fields_close(a_o, a_g)


SRC_CASES = {
    "plane_z_cw_tilted": dict(source="plane_z"),
    "plane_x_cw": dict(source="plane_x", shape=(16, 10, 12)),
    "plane_y_neg": dict(source="plane_y", shape=(10, 16, 12)),
    "pulse": dict(source="pulse"),
    "gated": dict(source="gated"),
    "table": dict(source="table"),
    "dipole": dict(source="dipole"),
    "plane_nonuniform": dict(source="plane_z", nonuniform=True, eps_tier=3),
    "plane_z_ragged": dict(source="plane_z", shape=(12, 10, 15)),
    "plane_x_ragged": dict(source="plane_x", shape=(16, 10, 13)),
    "plane_periodic": dict(source="plane_z", boundaries={"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}),
}


@pytest.mark.parametrize("name", list(SRC_CASES))
def test_source_injection_parity(name):
    objects, arrays, cfg = build_scene(**SRC_CASES[name], time=8e-15)
    steps = min(cfg.time_steps_total, 60)
    a_o, a_g = run_both(objects, arrays, cfg, steps, seed=False)
    assert np.abs(a_o.fields.E).max() > 0
    assert_fields_close(a_o, a_g)


ALL_DETS = ("field", "raw_field", "field_reduce", "energy", "energy_slices", "energy_pos", "energy_reduce",
            "poynting", "poynting_full", "poynting_all", "phasor", "phasor_pulse", "phasor_reduce")


@pytest.mark.parametrize("kw", [dict(), dict(nonuniform=True, eps_tier=3, mu_tier=3), dict(boundaries="periodic")])
def test_detector_parity(kw):
    objects, arrays, cfg = build_scene(source="plane_z", detectors=ALL_DETS, time=8e-15, **kw)
    steps = min(cfg.time_steps_total, 50)
    a_o, a_g = run_both(objects, arrays, cfg, steps, seed=False)
    assert_fields_close(a_o, a_g)
    assert_detectors_close(a_o, a_g)


@pytest.mark.parametrize(
    "modules",
    [
        [],
        [fx.DtypeConversion(dtype="bfloat16")],
        [fx.DtypeConversion(dtype="float16")],
        [fx.LinearReconstructEveryK(k=5), fx.DtypeConversion(dtype="float8_e4m3fnuz")],
        [fx.LinearReconstructEveryK(k=3)],
        [fx.DtypeConversion(dtype="float8_e4m3fn")],
        [fx.DtypeConversion(dtype="float8_e5m2")],
    ],
)
def test_forward_then_backward(modules):
    """C1-style: forward with boundary recording, then time-reversed steps with reset_fields=True and
    inverse detectors (backward.py:18-135).  The reverse pass through a CPML-terminated box amplifies
    rounding noise exponentially (psi is frozen), so the comparison runs a bounded number of reverse
    steps - like the reference's own 1- and 10-step reversal tests."""
    import torch

    rec = fx.Recorder(modules=modules)
    objects, arrays, cfg = build_scene(shape=(20, 18, 24), thickness=5, source="plane_z", detectors=("energy_slices", "inverse_energy"), recorder=rec, time=6e-15)
    T = cfg.time_steps_total
    st_o = yee.checkpointed_fdtd(arrays, objects, cfg)
    dev = arrays.to_torch("cuda")
    st_g = fx.run_fdtd(dev, objects, cfg)
    assert st_g[0] == T
    assert_fields_close(st_o[1], st_g[1])
    for k, ref in st_o[1].recording_state.data.items():
        got = st_g[1].recording_state.data[k]
        ref_f = ref.t.to(torch.float32).numpy() if hasattr(ref, "t") else ref
        # a 1-ulp float32 difference can flip a low-precision rounding: detector-level tolerance
        assert rel_l2(_to_np(got), ref_f) <= DET_TOL, k
    nback = 12
    st_o = yee.full_backward(st_o, objects, cfg, record_detectors=True, reset_fields=True, start_time_step=T - nback)
    st_g = fx.full_backward(st_g, objects, cfg, record_detectors=True, reset_fields=True, start_time_step=T - nback)
    assert st_g[0] == T - nback
    assert_fields_close(st_o[1], st_g[1], tol=DET_TOL)
    assert_detectors_close(st_o[1], st_g[1])


@pytest.mark.parametrize("kw", [dict(boundaries="periodic"), dict(boundaries="periodic", sigma_E=True, sigma_H=True, eps_tier=3, mu_tier=1),
                                dict(boundaries={"min_x": "pec", "max_x": "pec", "min_y": "pmc", "max_y": "pmc", "min_z": "periodic", "max_z": "periodic"})])
def test_reverse_undoes_forward(kw):
    """Size-independent property (tests/simulation/fdtd/test_time_reversal.py:251-327): one forward
    step followed by one reverse step reconstructs E and H (no PML => exactly reversible)."""
    import torch

    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build_scene(recorder=rec, source="dipole", **kw)
    seed_fields(arrays, seed=5)
    if any(isinstance(b, (fx.PerfectElectricConductor, fx.PerfectMagneticConductor)) for b in objects.boundary_objects):
        # start from a state that satisfies the wall conditions
        for b in objects.boundary_objects:
            comps = b.tangential_components if hasattr(b, "tangential_components") else ()
            f = arrays.fields.E if isinstance(b, fx.PerfectElectricConductor) else arrays.fields.H
            for c in comps:
                f[(c, *b.grid_slice)] = 0
    E0, H0 = arrays.fields.E.copy(), arrays.fields.H.copy()
    dev = arrays.to_torch("cuda")
    st = fx.forward((3, dev), cfg, objects, None, record_detectors=False, record_boundaries=True, simulate_boundaries=True)
    assert rel_l2(_to_np(st[1].fields.E), E0) > 1e-3
    st = fx.backward(st, cfg, objects, None, record_detectors=False, reset_fields=False)
    assert st[0] == 3
    assert np.allclose(_to_np(st[1].fields.E), E0, atol=1e-5 * np.abs(E0).max() * 10)
    assert np.allclose(_to_np(st[1].fields.H), H0, atol=1e-5 * np.abs(H0).max() * 10)


def test_large_grid_properties():
    """At a size the oracle cannot follow: vacuum box, periodic, no source - the discrete energy-like
    invariant of the leapfrog scheme stays bounded and a forward/reverse round trip returns the state."""
    import torch

    shape = (96, 160, 256)
    rec = fx.Recorder(modules=[])
    cfg = fx.SimulationConfig(time=40e-15, grid=fx.UniformGrid(spacing=50e-9), gradient_config=fx.GradientConfig(recorder=rec))
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = fx.boundary_objects_from_config(shape, cfg, "periodic")
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl], cfg, inv_permittivities=np.ones((1, *shape), np.float32))
    g = torch.Generator(device="cpu").manual_seed(0)
    dev = arrays.to_torch("cuda")
    dev.fields.E.copy_(1e-3 * torch.randn(dev.fields.E.shape, generator=g))
    dev.fields.H.copy_(1e-3 * torch.randn(dev.fields.H.shape, generator=g))
    E0, H0 = dev.fields.E.clone(), dev.fields.H.clone()
    n = 20
    st = (0, dev)
    from fdtdx_b200.fdtd import get_plan

    plan = get_plan(dev, objects, cfg)
    plan.run_forward(0, n, False, True, True)
    e1 = float((dev.fields.E.double() ** 2).sum() + (dev.fields.H.double() ** 2).sum())
    e0 = float((E0.double() ** 2).sum() + (H0.double() ** 2).sum())
    assert 0.2 * e0 < e1 < 5.0 * e0  # white-noise start: the leapfrog pseudo-energy, not E^2+H^2, is conserved
    plan.run_reverse(n, n, False, False)
    assert float((dev.fields.E - E0).abs().max()) < 1e-6
    assert float((dev.fields.H - H0).abs().max()) < 1e-6


def test_errors_are_loud():
    import torch

    objects, arrays, cfg = build_scene()
    with pytest.raises(RuntimeError):
        fx.run_fdtd(arrays, objects, cfg)  # numpy container: no CPU fallback
    objects, arrays, cfg = build_scene(poles=1, recorder=fx.Recorder(modules=[]))
    dev = arrays.to_torch("cuda")
    with pytest.raises(NotImplementedError):
        fx.reversible_fdtd(dev, objects, cfg)
    with pytest.raises(NotImplementedError):
        fx.backward((3, dev), cfg, objects)
