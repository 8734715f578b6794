"""config.symmetry on the hot path: the one-sided halo rule of ``fdtd/update.py:121-125`` and
the detector mirror ``pad_fields_with_symmetry_mirror`` (``update.py:139-198``).  The reduced half-domain itself
is built at setup time (``fdtd/symmetry.py``, out of scope); here the wall objects are placed by hand.  CPU tests
re-type the reference's known answers (``tests/unit/fdtd/test_update.py:190-267``) against the oracle; ``-m gpu``
tests compare the CUDA kernels with the oracle on scenes whose detectors touch the symmetry plane."""

import dataclasses

import numpy as np
import pytest

import fdtdx_b200 as fx
from oracle import yee
from scenes import make_config, rel_l2, seed_fields

F = np.float32


def build(shape, types, symmetry, sym_wall_axes=(), drop=(), source=True, detectors=("energy", "poynting", "field"), time=6e-15, thickness=3):
    cfg = make_config(shape, time=time)
    cfg = cfg.aset("symmetry", tuple(symmetry))
    nx, ny, nz = shape
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=((0, nx), (0, ny), (0, nz)))
    bl = [b for b in fx.boundary_objects_from_config(shape, cfg, types, thickness=thickness) if not any(b.name.endswith(d) for d in drop)]
    for b in bl:
        if isinstance(b, fx.PerfectElectricConductor) and b.axis in sym_wall_axes and b.direction == "-":
            b._is_symmetry_wall = True
    rng = np.random.default_rng(5)
    inv_eps = (1.0 / (1.0 + 3.0 * rng.random((1, *shape)))).astype(F)
    objs = [vol, *bl]
    wc = fx.WaveCharacter(wavelength=0.8e-6)
    if source:
        objs.append(fx.make_plane_source("source", ((0, nx), (0, ny), (thickness + 1, thickness + 2)), cfg, inv_eps, 1.0, direction="+", wave_character=wc,
                                         fixed_E_polarization_vector=(1.0, 0.0, 0.0)))
    full = ((0, nx), (0, ny), (0, nz))
    mk = {
        "energy": lambda: fx.EnergyDetector(name="energy", grid_slice_tuple=full, as_slices=True),
        "poynting": lambda: fx.PoyntingFluxDetector(name="poynting", grid_slice_tuple=((0, nx), (0, ny), (nz - thickness - 3, nz - thickness - 2)), direction="+"),
        "field": lambda: fx.FieldDetector(name="field", grid_slice_tuple=((0, nx), (0, 3), (2, nz - 2))),
    }
    objs += [mk[d]() for d in detectors]
    objects, arrays, _, cfg, _ = fx.place_objects(objs, cfg, inv_permittivities=inv_eps)
    return objects, arrays, cfg


# ------------------------------------------------------------------------------------------ CPU: known answers
def _mock_objects(shape, axis, wrap, is_symmetry_wall):
    cfg = make_config(shape)
    b = fx.BlochBoundary(name="b", grid_slice_tuple=tuple((0, n) for n in shape), axis=axis, direction="+") if wrap else \
        fx.PerfectElectricConductor(name="b", grid_slice_tuple=tuple((0, n) for n in shape), axis=axis, direction="-")
    b._is_symmetry_wall = is_symmetry_wall
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=tuple((0, n) for n in shape))
    return fx.ObjectContainer(object_list=[vol, b], volume_idx=0), cfg


def test_symmetric_axis_does_not_wrap_its_min_side_halo():
    """test_update.py:190-208."""
    shape = (2, 3, 2)
    fields = np.arange(1, 3 * 12 + 1, dtype=F).reshape(3, *shape)
    objects, cfg = _mock_objects(shape, 1, wrap=True, is_symmetry_wall=False)
    wrapped = yee.pad_fields_for_boundaries(fields, objects, cfg)
    symmetric = yee.pad_fields_for_boundaries(fields, objects, cfg.aset("symmetry", (0, 1, 0)))
    assert np.allclose(wrapped[:, 1:-1, 0, 1:-1], fields[:, :, -1, :])
    assert np.all(symmetric[:, :, 0, :] == 0.0)
    assert np.allclose(symmetric[:, 1:-1, -1, 1:-1], fields[:, :, 0, :])


def test_electric_plane_mirrors_each_component_with_its_own_index_map():
    """test_update.py:236-254, :256-267: E tangential (odd, on the plane) = minus the SECOND cell, E normal (even,
    half a cell off) = plus the first; H the other way round; magnetic planes and user walls keep the zero halo."""
    shape = (2, 3, 2)
    fields = np.arange(1, 3 * 12 + 1, dtype=F).reshape(3, *shape)
    objects, cfg = _mock_objects(shape, 1, wrap=False, is_symmetry_wall=True)
    cfg_e = cfg.aset("symmetry", (0, -1, 0))
    p = yee.pad_fields_with_symmetry_mirror(fields, objects, cfg_e, "E")
    assert np.allclose(p[0, 1:-1, 0, 1:-1], -fields[0, :, 1, :]) and np.allclose(p[1, 1:-1, 0, 1:-1], fields[1, :, 0, :]) and np.allclose(p[2, 1:-1, 0, 1:-1], -fields[2, :, 1, :])
    p = yee.pad_fields_with_symmetry_mirror(fields, objects, cfg_e, "H")
    assert np.allclose(p[0, 1:-1, 0, 1:-1], fields[0, :, 0, :]) and np.allclose(p[1, 1:-1, 0, 1:-1], -fields[1, :, 1, :]) and np.allclose(p[2, 1:-1, 0, 1:-1], fields[2, :, 0, :])
    for ft in ("E", "H"):
        assert np.all(yee.pad_fields_with_symmetry_mirror(fields, objects, cfg.aset("symmetry", (0, 1, 0)), ft)[:, :, 0, :] == 0.0)
    user, _ = _mock_objects(shape, 1, wrap=False, is_symmetry_wall=False)
    assert np.all(yee.pad_fields_with_symmetry_mirror(fields, user, cfg_e, "E")[:, :, 0, :] == 0.0)


# ------------------------------------------------------------------------------------------ GPU: CUDA vs oracle
CASES = {
    # electric symmetry wall on min-y (PEC object flagged as symmetry wall), PML elsewhere
    "electric_y": dict(types={"min_x": "pml", "max_x": "pml", "min_y": "pec", "max_y": "pml", "min_z": "pml", "max_z": "pml"}, symmetry=(0, -1, 0), sym_wall_axes=(1,)),
    # electric wall on min-x and min-y (corner: both mirrors), periodic far side on x (one-sided wrap)
    "electric_xy_far_periodic": dict(types={"min_x": "pec", "max_x": "periodic", "min_y": "pec", "max_y": "pml", "min_z": "pml", "max_z": "pml"}, symmetry=(-1, -1, 0), sym_wall_axes=(0, 1)),
    # magnetic symmetry on x: no wall object on the min side, far side still periodic -> min halo zero, max halo wraps
    "magnetic_x_far_periodic": dict(types={"min_x": "periodic", "max_x": "periodic", "min_y": "pml", "max_y": "pml", "min_z": "pml", "max_z": "pml"}, symmetry=(1, 0, 0), drop=("min_x",)),
    # magnetic symmetry on z with a periodic far side (lane-0 wrap site of the marching kernels)
    "magnetic_z_far_periodic": dict(types={"min_x": "pml", "max_x": "pml", "min_y": "pml", "max_y": "pml", "min_z": "periodic", "max_z": "periodic"}, symmetry=(0, 0, 1), drop=("min_z",), source=False),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_symmetry_halo_rules_match_oracle(case):
    kw = dict(CASES[case])
    objects, arrays, cfg = build((10, 8, 16), **kw)
    seed_fields(arrays, seed=2)
    T = min(10, cfg.time_steps_total)
    st_o = (0, arrays.map_arrays(lambda a: a.copy() if isinstance(a, np.ndarray) else a))
    for _ in range(T):
        st_o = yee.forward(st_o, cfg, objects, record_detectors=True)
    t, out = fx.custom_fdtd_forward(arrays.to_torch("cuda"), objects, cfg, reset_container=False, record_detectors=True, start_time=0, end_time=T)
    for name in ("E", "H"):
        e = rel_l2(getattr(out.fields, name).cpu().numpy(), getattr(st_o[1].fields, name))
        assert e <= 1e-5, (case, name, e)
    for d, st in st_o[1].detector_states.items():
        for key, ref in st.items():
            e = rel_l2(out.detector_states[d][key].cpu().numpy(), ref)
            assert e <= 1e-4 and np.abs(ref).max() > 0, (case, d, key, e)


@pytest.mark.gpu
def test_detector_mirror_changes_the_plane_row():
    """The mirror matters: without the symmetry-wall flag the detector row on the plane reads a zero halo instead."""
    kw = dict(CASES["electric_y"])
    outs = []
    for flagged in (True, False):
        objects, arrays, cfg = build((10, 8, 16), **dict(kw, sym_wall_axes=(1,) if flagged else ()))
        seed_fields(arrays, seed=2)
        t, out = fx.custom_fdtd_forward(arrays.to_torch("cuda"), objects, cfg, reset_container=False, record_detectors=True, start_time=0, end_time=4)
        outs.append(out.detector_states["field"]["fields"].cpu().numpy())
    assert np.abs(outs[0][:, :, :, 0] - outs[1][:, :, :, 0]).max() > 0     # the y = 0 row sees the mirror partner
    assert np.array_equal(outs[0][:, :, :, 1:], outs[1][:, :, :, 1:])       # every other row is untouched
