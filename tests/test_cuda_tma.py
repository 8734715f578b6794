"""TMA-staged half-step kernels (csrc/yee_tma.cuh) vs the register-marching kernels vs the oracle.

The two CUDA paths share the arithmetic (same float32 op order, no FMA contraction), so they must
agree BIT-EXACTLY on every scene the staged path accepts; the oracle comparison uses the north-star
tolerances.  The scenes stress what is specific to the staged path: tiles that overhang the grid
(Ny % 8 != 0, Nz % 128 != 0, several z tiles), x chunks shorter than the ring depth, chunk seams,
x wrap (periodic x with CPML on y/z), the neighbour-plane stage at the domain end, odd-thickness
CPML (scalar z-slab path), PEC/PMC faces (TMA out-of-bounds zero fill = the zero halo of
``pad_fields``, core/misc.py:615-641), sources on chunk seams and the reverse pass.
"""

import numpy as np
import pytest

import fdtdx_b200 as fx
from oracle import yee
from scenes import build_scene, rel_l2, seed_fields

pytestmark = pytest.mark.gpu

X_PERIODIC = {"min_x": "periodic", "max_x": "periodic", "min_y": "pml", "max_y": "pml", "min_z": "pml", "max_z": "pml"}
PEC_PMC = {"min_x": "pec", "max_x": "pmc", "min_y": "pmc", "max_y": "pec", "min_z": "pec", "max_z": "pml"}

SCENES = {
    # name: (build_scene kwargs, steps)
    "two_z_tiles_overhang": (dict(shape=(11, 21, 132), thickness=4), 5),
    "three_z_tiles": (dict(shape=(5, 9, 260), thickness=2), 5),
    "odd_thickness": (dict(shape=(10, 13, 20), thickness=3), 6),
    "kappa": (dict(shape=(10, 13, 20), thickness=4, kappa=True), 5),
    "x_periodic": (dict(shape=(9, 12, 24), thickness=4, boundaries=X_PERIODIC), 6),
    "pec_pmc": (dict(shape=(9, 11, 16), thickness=3, boundaries=PEC_PMC), 6),
    "nonuniform_all": (dict(shape=(10, 12, 16), nonuniform=True, eps_tier=3, sigma_E=True, mu_tier=3, sigma_H=True), 5),
    "mu_iso": (dict(shape=(8, 9, 12), mu_tier=1), 4),
    "ade": (dict(shape=(8, 10, 16), poles=2, c4=True, sigma_E=True, eps_tier=3, coeff_tier=3), 5),
}


def _np(x):
    return x.detach().cpu().numpy()


def _run(objects, arrays, cfg, steps, tma, xchunk, record_detectors=False):
    from fdtdx_b200.fdtd import get_plan

    dev = arrays.to_torch("cuda")
    objects.__dict__.pop("_plan_cache", None)
    plan = get_plan(dev, objects, cfg)
    plan.set_tma(tma, xchunk)
    plan.run_forward(0, steps, record_detectors, False, True)
    out = plan.finish(dev)
    objects.__dict__.pop("_plan_cache", None)
    return out


def _assert_identical(a, b):
    assert np.array_equal(_np(a.fields.E), _np(b.fields.E)), "E differs between the two CUDA paths"
    assert np.array_equal(_np(a.fields.H), _np(b.fields.H)), "H differs between the two CUDA paths"
    for name in a.fields.psi_E:
        for w in range(2):
            assert np.array_equal(_np(a.fields.psi_E[name][w]), _np(b.fields.psi_E[name][w])), f"psi_E {name}"
            assert np.array_equal(_np(a.fields.psi_H[name][w]), _np(b.fields.psi_H[name][w])), f"psi_H {name}"
    if a.fields.dispersive_P_curr is not None:
        assert np.array_equal(_np(a.fields.dispersive_P_curr), _np(b.fields.dispersive_P_curr))


@pytest.mark.parametrize("xchunk", [1, 2, 5, 0])
@pytest.mark.parametrize("name", list(SCENES))
def test_staged_equals_marching_and_oracle(name, xchunk):
    kw, steps = SCENES[name]
    objects, arrays, cfg = build_scene(**kw)
    seed_fields(arrays, seed=5)
    ref = _run(objects, arrays, cfg, steps, tma=0, xchunk=0)
    got = _run(objects, arrays, cfg, steps, tma=1, xchunk=xchunk)
    _assert_identical(ref, got)
    if xchunk == 0:
        st = (0, arrays)
        for _ in range(steps):
            st = yee.forward(st, cfg, objects, None, False, False, True)
        assert rel_l2(_np(got.fields.E), st[1].fields.E) <= 1e-5
        assert rel_l2(_np(got.fields.H), st[1].fields.H) <= 1e-5


@pytest.mark.parametrize("src", ["plane_x", "plane_z", "dipole", "gated"])
@pytest.mark.parametrize("xchunk", [3, 0])
def test_staged_sources_and_detectors(src, xchunk):
    """Source planes land on chunk seams for xchunk=3; detectors read the staged kernels' output."""
    shape = (16, 10, 12) if src == "plane_x" else (12, 10, 16)
    objects, arrays, cfg = build_scene(shape=shape, source=src, detectors=("energy_slices", "phasor", "poynting"), time=6e-15)
    steps = min(cfg.time_steps_total, 40)
    ref = _run(objects, arrays, cfg, steps, tma=0, xchunk=0, record_detectors=True)
    got = _run(objects, arrays, cfg, steps, tma=1, xchunk=xchunk, record_detectors=True)
    assert float(np.abs(_np(got.fields.E)).max()) > 0
    _assert_identical(ref, got)
    for name, st in ref.detector_states.items():
        for key, v in st.items():
            assert np.array_equal(_np(v), _np(got.detector_states[name][key])), (name, key)


def test_staged_reverse_pass():
    """reversible pass: forward with boundary recording, then time-reversed steps on the staged
    kernels (update.py:526-609, 856-930) reproduce the oracle's reconstruction."""
    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build_scene(shape=(14, 12, 16), thickness=4, eps_tier=3, source="plane_z", recorder=rec, time=4e-15)
    T = cfg.time_steps_total
    st_o = yee.checkpointed_fdtd(arrays, objects, cfg)
    st_o = yee.full_backward(st_o, objects, cfg, record_detectors=False, reset_fields=True, start_time_step=T - 7)
    outs = []
    for tma in (0, 1):
        import os

        os.environ["FDTDX_B200_TMA"] = str(tma)
        try:
            objects.__dict__.pop("_plan_cache", None)
            st_g = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg)
            st_g = fx.full_backward(st_g, objects, cfg, record_detectors=False, reset_fields=True, start_time_step=T - 7)
            outs.append(st_g[1])
        finally:
            os.environ.pop("FDTDX_B200_TMA", None)
    assert rel_l2(_np(outs[1].fields.E), st_o[1].fields.E) <= 1e-4
    assert rel_l2(_np(outs[1].fields.H), st_o[1].fields.H) <= 1e-4
    assert rel_l2(_np(outs[1].fields.E), _np(outs[0].fields.E)) <= 1e-6


def test_property_large_grid_paths_agree():
    """Full-size-ish property check: 3 steps on 40x100x256 (two z tiles, 13 row tiles), CPML + metric."""
    objects, arrays, cfg = build_scene(shape=(40, 100, 256), thickness=6, nonuniform=True)
    seed_fields(arrays, seed=11)
    ref = _run(objects, arrays, cfg, 3, tma=0, xchunk=0)
    got = _run(objects, arrays, cfg, 3, tma=1, xchunk=0)
    _assert_identical(ref, got)


def _random_case(seed):
    """Seeded random scene: shape, per-face boundary kinds, CPML thickness, material tiers, metric."""
    rng = np.random.default_rng(1000 + seed)
    nz = int(rng.choice([4, 8, 12, 20, 36, 132, 140]))
    ny = int(rng.integers(1, 20))
    nx = int(rng.integers(1, 14))
    th = int(rng.integers(1, 5))
    kinds = {}
    for ax, n in zip("xyz", (nx, ny, nz)):
        if ax == "x" and rng.random() < 0.3:
            kinds[f"min_{ax}"] = kinds[f"max_{ax}"] = "periodic"
            continue
        for side in ("min", "max"):
            kinds[f"{side}_{ax}"] = str(rng.choice(["pml", "pec", "pmc"])) if n >= 2 * th + 2 else str(rng.choice(["pec", "pmc"]))
    kw = dict(shape=(nx, ny, nz), thickness=th, boundaries=kinds,
              eps_tier=int(rng.choice([1, 3])), mu_tier=int(rng.choice([0, 1, 3])),
              sigma_E=bool(rng.random() < 0.3), sigma_H=bool(rng.random() < 0.3),
              nonuniform=bool(rng.random() < 0.5), kappa=bool(rng.random() < 0.3))
    return kw, int(rng.integers(2, 6)), int(rng.choice([0, 1, 2, 3, 7]))


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_staged_equals_marching_and_oracle(seed):
    """Randomised scenes (degenerate extents included: Nx or Ny of 1, slabs touching, 2 z tiles)."""
    kw, steps, xchunk = _random_case(seed)
    objects, arrays, cfg = build_scene(**kw)
    seed_fields(arrays, seed=seed)
    ref = _run(objects, arrays, cfg, steps, tma=0, xchunk=0)
    got = _run(objects, arrays, cfg, steps, tma=1, xchunk=xchunk)
    _assert_identical(ref, got)
    st = (0, arrays)
    for _ in range(steps):
        st = yee.forward(st, cfg, objects, None, False, False, True)
    scale = max(float(np.abs(st[1].fields.E).max()), 1e-30)
    assert rel_l2(_np(got.fields.E), st[1].fields.E) <= 1e-5, (kw, float(scale))
    assert rel_l2(_np(got.fields.H), st[1].fields.H) <= 1e-5, kw


def test_full_size_coupler_paths_agree():
    """BASELINE configs[1] at full size (1897x291x128, 70.7 Mcell, metric + CPML + mode source + phasor
    detectors): the staged and the marching kernels produce identical fields and detector states."""
    import torch
    from fdtdx_b200 import workloads as W
    from fdtdx_b200.fdtd import get_plan

    outs = []
    for tma in (0, 1):
        objects, arrays, cfg = W.build_coupler(20, device=torch.device("cuda"))
        torch.manual_seed(0)
        arrays.fields.E.copy_(1e-3 * torch.randn_like(arrays.fields.E))
        arrays.fields.H.copy_(1e-3 * torch.randn_like(arrays.fields.H))
        plan = get_plan(arrays, objects, cfg)
        plan.set_tma(tma, 0)
        plan.run_forward(0, 3, True, False, True)
        torch.cuda.synchronize()
        outs.append((arrays.fields.E.clone(), arrays.fields.H.clone(), {k: {k2: v2.clone() for k2, v2 in v.items()} for k, v in arrays.detector_states.items()}))
        del plan, arrays, objects
        torch.cuda.empty_cache()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert float(outs[1][0].abs().max()) > 0
    for k, v in outs[0][2].items():
        for k2, v2 in v.items():
            assert torch.equal(v2, outs[1][2][k][k2]), (k, k2)
