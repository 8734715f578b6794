"""Unfold helpers of symmetry-reduced runs (fdtdx_b200/symmetry.py) against the REFERENCE'S OWN SOURCE: its
``fdtd/symmetry.py`` and ``core/physics/symmetry.py`` are executed from /root/reference under the NumPy ``jax.numpy``
stand-in (oracle/refexec.py) and every function is compared bit for bit on seeded arrays.  Where the reference tree is
absent (the GPU box) the same cases are checked against closed-form expectations re-typed from its docstrings."""

import itertools

import numpy as np
import pytest

import fdtdx_b200 as fx
from fdtdx_b200 import symmetry as S
from oracle import refexec

SYMS = [s for s in itertools.product((-1, 0, 1), repeat=3) if any(s)]


@pytest.fixture(scope="module")
def ref():
    if not refexec.available():
        pytest.skip("reference sources not present")
    refexec.Reference.FILES = dict(refexec.Reference.FILES, **{"fdtdx.fdtd.symmetry": "fdtd/symmetry.py"})
    r = refexec.Reference()
    # objects/detectors/diffractive.py is not among the executed files: its import is a catch-all stub that every object
    # "is an instance of"; give the isinstance test in _unfold_one_detector a class nothing here derives from
    r.modules["fdtdx.fdtd.symmetry"].DiffractiveDetector = type("DiffractiveDetector", (), {})
    return r


def _rand(shape, seed, cplx=False):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(shape).astype(np.float32)
    return (a + 1j * rng.standard_normal(shape).astype(np.float32)).astype(np.complex64) if cplx else a


def test_parity_and_index_map_tables_match_reference(ref):
    m = ref.modules["fdtdx.core.physics.symmetry"]
    for ft, c, a, w in itertools.product("EH", range(3), range(3), (-1, 1)):
        assert S.field_component_parity(ft, c, a, w) == m.field_component_parity(ft, c, a, w)
        assert S.mirror_pairs_on_plane(ft, c, a, w) == m.mirror_pairs_on_plane(ft, c, a, w)
        assert S.component_sits_on_plane(ft, c, a) == m.component_sits_on_plane(ft, c, a)
    x = _rand((2, 5, 4), 0)
    for axis, parity, on in itertools.product(range(3), (-1, 1), (False, True)):
        assert np.array_equal(S.mirror_extend_low_side(x, axis, parity, on), np.asarray(m.mirror_extend_low_side(ref.jnp.asarray(x), axis, parity, on)))


@pytest.mark.parametrize("sym", SYMS)
def test_unfold_fields_and_array_match_reference(ref, sym):
    m = ref.modules["fdtdx.fdtd.symmetry"]
    f = _rand((3, 4, 5, 6), 1)
    for ft in "EH":
        want = np.asarray(m.unfold_fields(ref.jnp.asarray(f), sym, ft))
        got = S.unfold_fields(f, sym, ft)
        assert got.shape == tuple(n * (2 if s else 1) for n, s in zip((4, 5, 6), sym))[0:0] + want.shape and np.array_equal(got, want)
    arr = _rand((2, 3, 4, 5, 6), 2)
    signs = {a: np.asarray([1.0, -1.0, 1.0], np.float32).reshape(1, 3, 1, 1, 1) for a in range(3) if sym[a]}
    for on in ((), (0,), (0, 1)):
        want = np.asarray(m.unfold_array(ref.jnp.asarray(arr), sym, (2, 3, 4), {a: ref.jnp.asarray(v) for a, v in signs.items()}, on))
        assert np.array_equal(S.unfold_array(arr, sym, (2, 3, 4), signs, on), want)


def _ref_detector(ref, kind, **attrs):
    """An instance of the reference's detector class carrying just the attributes _unfold_one_detector reads."""
    mod = {"PhasorDetector": "phasor", "FieldDetector": "field", "EnergyDetector": "energy", "PoyntingFluxDetector": "poynting_flux"}[kind]
    cls = getattr(ref.modules[f"fdtdx.objects.detectors.{mod}"], kind)
    shim = type("Shim" + kind, (cls,), {"propagation_axis": attrs.pop("propagation_axis", 2), "__init__": lambda self: None, "__setattr__": object.__setattr__})
    obj = shim()
    for k, v in attrs.items():
        object.__setattr__(obj, k, v)
    return obj


DET_CASES = {
    "phasor_spatial": ("PhasorDetector", dict(components=("Ex", "Ez", "Hy"), reduce_volume=False, exact_interpolation=True), {"phasor": ((1, 2, 3, 4, 5, 6), True)}),
    "phasor_reduced": ("PhasorDetector", dict(components=("Ex", "Hy", "Hz"), reduce_volume=True, exact_interpolation=True), {"phasor": ((1, 2, 3), True)}),
    "field_raw": ("FieldDetector", dict(components=("Ey", "Hx"), reduce_volume=False, exact_interpolation=False), {"fields": ((4, 2, 4, 5, 6), False)}),
    "field_reduced": ("FieldDetector", dict(components=("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"), reduce_volume=True, exact_interpolation=True), {"fields": ((4, 6), False)}),
    "energy_slices": ("EnergyDetector", dict(as_slices=True, reduce_volume=False, exact_interpolation=True),
                      {"XY Plane": ((3, 4, 5), False), "XZ Plane": ((3, 4, 6), False), "YZ Plane": ((3, 5, 6), False)}),
    "energy_volume": ("EnergyDetector", dict(as_slices=False, reduce_volume=False, exact_interpolation=True), {"energy": ((3, 4, 5, 6), False)}),
    "energy_reduced": ("EnergyDetector", dict(as_slices=False, reduce_volume=True, exact_interpolation=True), {"energy": ((3, 1), False)}),
    "poynting_plane": ("PoyntingFluxDetector", dict(reduce_volume=False, keep_all_components=False, exact_interpolation=True, propagation_axis=2), {"poynting_flux": ((3, 4, 5, 1), False)}),
    "poynting_all": ("PoyntingFluxDetector", dict(reduce_volume=False, keep_all_components=True, exact_interpolation=True, propagation_axis=2), {"poynting_flux": ((3, 3, 4, 5, 6), False)}),
    "poynting_reduced_all": ("PoyntingFluxDetector", dict(reduce_volume=True, keep_all_components=True, exact_interpolation=True, propagation_axis=1), {"poynting_flux": ((3, 3), False)}),
    "poynting_reduced": ("PoyntingFluxDetector", dict(reduce_volume=True, keep_all_components=False, exact_interpolation=True, propagation_axis=0), {"poynting_flux": ((3, 1), False)}),
}


def _host_detector(kind, attrs):
    box = ((0, 4), (0, 5), (0, 6))
    a = dict(attrs)
    a.pop("propagation_axis", None)
    if kind == "PhasorDetector":
        return fx.PhasorDetector(name="d", grid_slice_tuple=box, wave_characters=(fx.WaveCharacter(wavelength=1e-6), fx.WaveCharacter(wavelength=2e-6)), **a)
    if kind == "PoyntingFluxDetector":
        return fx.PoyntingFluxDetector(name="d", grid_slice_tuple=box, direction="+", fixed_propagation_axis=attrs["propagation_axis"], **a)
    return getattr(fx, kind)(name="d", grid_slice_tuple=box, **a)


@pytest.mark.parametrize("touched", [(-1, 0, 0), (1, -1, 0), (-1, -1, 1), (0, 0, 1)])
@pytest.mark.parametrize("name", list(DET_CASES))
def test_unfold_one_detector_matches_reference(ref, name, touched):
    m = ref.modules["fdtdx.fdtd.symmetry"]
    kind, attrs, leaves = DET_CASES[name]
    state = {k: _rand(shape, 3 + i, cplx) for i, (k, (shape, cplx)) in enumerate(leaves.items())}
    count = sum(1 for t in touched if t)
    want = m._unfold_one_detector(_ref_detector(ref, kind, **attrs), {k: ref.jnp.asarray(v) for k, v in state.items()}, touched, count)
    got = S._unfold_one_detector(_host_detector(kind, attrs), state, touched, count)
    assert set(got) == set(want)
    for k in want:
        w = np.asarray(want[k])
        assert got[k].shape == w.shape, (k, got[k].shape, w.shape)
        assert np.array_equal(np.asarray(got[k], dtype=w.dtype), w), k


def test_unfold_closed_forms():
    """Known answers that need no reference tree: an even, half-cell-offset component doubles by a plain flip; an on-plane
    component keeps index 0 as its own mirror; odd reductions vanish, even ones double (sum) or stay (mean)."""
    f = np.zeros((3, 3, 1, 1), np.float32)
    f[:, :, 0, 0] = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    e = S.unfold_fields(f, (-1, 0, 0), "E")[:, :, 0, 0]  # electric plane normal to x: Ex even and off the plane, Ey / Ez odd and on it
    assert e[0].tolist() == [3, 2, 1, 1, 2, 3]
    assert e[1].tolist() == [-6, -6, -5, 4, 5, 6]
    h = S.unfold_fields(f, (1, 0, 0), "H")[:, :, 0, 0]  # magnetic plane: every component mirrors one-to-one; Hx even, Hy / Hz odd
    assert h[0].tolist() == [3, 2, 1, 1, 2, 3] and h[2].tolist() == [-9, -8, -7, 7, 8, 9]
    det = fx.FieldDetector(name="d", grid_slice_tuple=((0, 2), (0, 2), (0, 2)), components=("Ex", "Ey"), reduce_volume=True)
    out = S._unfold_one_detector(det, {"fields": np.ones((2, 2), np.float32)}, (-1, 0, 0), 1)["fields"]
    assert out[:, 0].tolist() == [1.0, 1.0] and out[:, 1].tolist() == [0.0, 0.0]  # mean: even kept, odd vanishes
    with pytest.raises(ValueError):
        S.unfold_fields(f, (0, 0, 0), "E")


def test_unfold_detector_states_on_torch_tensors():
    import torch

    box = ((0, 4), (0, 5), (2, 6))
    dets = [fx.EnergyDetector(name="e", grid_slice_tuple=box, reduce_volume=True), fx.FieldDetector(name="f", grid_slice_tuple=((1, 4), (0, 5), (2, 6)), components=("Ez",))]

    class _Arrays:
        def __init__(self, st):
            self.detector_states = st

        def aset(self, key, value):
            assert key == "detector_states"
            return _Arrays(value)

    class _Objects:
        detectors = dets

    class _Cfg:
        symmetry = (-1, 1, 0)

    st = {"e": {"energy": torch.ones(3, 1)}, "f": {"fields": torch.arange(2 * 1 * 3 * 5 * 4, dtype=torch.float32).reshape(2, 1, 3, 5, 4)}}
    out = S.unfold_detector_states(_Arrays(st), _Objects(), _Cfg()).detector_states
    assert torch.equal(out["e"]["energy"], 4 * torch.ones(3, 1))          # both planes clip it: x 2^2
    assert out["f"]["fields"].shape == (2, 1, 3, 10, 4)                    # starts at x = 1: only the y plane clips it
    assert torch.equal(out["f"]["fields"][:, :, :, 5:], st["f"]["fields"])
    assert torch.equal(out["f"]["fields"][:, :, :, :5], st["f"]["fields"].flip(3))  # Ez is even across a magnetic plane normal to y
