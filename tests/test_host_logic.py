"""Host-side logic and the C-ABI surface (no GPU needed): switches, slab partitioning, plan-table
helpers, and that the built library loads and exports every symbol include/fdtdx_b200.h declares."""

import ctypes
import os
import re

import numpy as np
import pytest

import fdtdx_b200 as fx
from fdtdx_b200 import _lib
from fdtdx_b200.dist import neighbours, slab_bounds
from fdtdx_b200.plan import metric_scales
from fdtdx_b200.workloads import build_box, build_coupler, bytes_per_cell_step, coupler_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "fdtdx_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(fdtdx_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built (run __graft_entry__.build())")
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) <= set(declared)
    assert L.fdtdx_b200_version() >= 100


def test_plan_creation_errors_without_gpu_are_reported():
    """Argument validation happens before any CUDA call, so it can be checked on a CPU box."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.fdtdx_b200_plan_create(ctypes.byref(h), 0, 4, 4, 0, 4, 0.5, 1e-16, 10, 1, 0, 0, 0, 1.0, None, None, None, None)
    assert rc == -1 and b"bad dimensions" in L.fdtdx_b200_last_error()
    rc = L.fdtdx_b200_plan_create(ctypes.byref(h), 4, 4, 4, 0, 4, 0.5, 1e-16, 10, 2, 0, 0, 0, 1.0, None, None, None, None)
    assert rc == -1 and b"eps_tier" in L.fdtdx_b200_last_error()


def test_no_cpu_fallback():
    from scenes import build_scene

    objects, arrays, cfg = build_scene()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        fx.run_fdtd(arrays, objects, cfg)


def test_switch_tables():
    """core/switch.py semantics: interval, start/end, fixed steps."""
    dt = 1e-16
    assert fx.OnOffSwitch().is_default_always_on
    assert fx.OnOffSwitch(interval=3).calculate_on_list(7, dt) == [True, False, False, True, False, False, True]
    on = fx.OnOffSwitch(start_time=2e-16, end_time=4e-16).calculate_on_list(6, dt)
    assert on == [False, False, True, True, True, False]
    idx = fx.OnOffSwitch(fixed_on_time_steps=[1, 4]).calculate_time_step_to_on_arr_idx(5, dt)
    assert idx == [-1, 0, -1, -1, 1]
    with pytest.raises(Exception):
        fx.OnOffSwitch(start_after_periods=1.0).is_on_at_time_step(0, dt)
    assert fx.OnOffSwitch(is_always_off=True).calculate_on_list(3, dt) == [False] * 3


def test_config_courant_and_steps():
    cfg = fx.SimulationConfig(time=100e-15, grid=fx.UniformGrid(spacing=100e-9))
    assert cfg.time_steps_total == 525  # examples/simulate_gaussian_source.py (C1)
    assert abs(cfg.courant_number - 0.99 / np.sqrt(3)) < 1e-12
    g, nx1, _ = coupler_grid(20)
    assert g.shape == (1897, 291, 128) and nx1 == 1897  # performance/directional_coupler.py at cpl=20
    assert coupler_grid(10)[0].shape == (948, 145, 65)
    cfg = fx.SimulationConfig(time=2.0 * 3.5 * 42e-6 / 3e8, grid=g)
    assert cfg.has_nonuniform_grid and abs(cfg.time_steps_total - 23188) <= 2


def test_metric_scales_match_reference_definition():
    """curl.py:29-39: backward = ref / mean(w_i, w_{i-1}) with w_{-1} := w_0; forward = ref / w_i."""
    g = fx.RectilinearGrid([0, 1, 3, 6], [0, 1, 2, 3], [0, 2, 4, 6])
    cfg = fx.SimulationConfig(time=1e-15, grid=g)
    ref = fx.constants.c * cfg.time_step_duration / cfg.courant_number
    sB, sF = metric_scales(cfg, 0)
    assert np.allclose(sF, ref / np.array([1, 2, 3.0]), rtol=1e-6)
    assert np.allclose(sB, ref / np.array([1, 1.5, 2.5]), rtol=1e-6)


def test_slab_partition():
    assert slab_bounds(1600, 8, 3) == (600, 800)
    with pytest.raises(ValueError):
        slab_bounds(10, 4, 0)
    assert neighbours(0, 4, False) == (None, 1) and neighbours(3, 4, False) == (2, None)
    assert neighbours(0, 4, True) == (3, 1) and neighbours(3, 4, True) == (2, 0)
    assert neighbours(0, 1, True) == (None, None)


def test_workload_builders_and_byte_model():
    objects, arrays, cfg = build_box((32, 32, 32), device=None)
    assert arrays.fields.E.shape == (3, 32, 32, 32) and len(objects.pml_objects) == 6
    b = bytes_per_cell_step(objects, arrays, (32, 32, 32))
    psi_cells = 6 * 10 * 32 * 32
    assert abs(b - (76 + 32 * psi_cells / 32**3)) < 1e-9  # 76 B/cell + 32 B per CPML membership
    objects, arrays, cfg = build_coupler(4, device=None)
    assert len(objects.detectors) == 3 and len(objects.sources) == 1
    assert abs(1 / float(arrays.inv_permittivities.min()) - 12.25) < 1e-5 and abs(1 / float(arrays.inv_permittivities.max()) - 2.25) < 1e-5
    # x-slab allocation only holds the slab and the PML slabs that intersect it
    nx = objects.volume.grid_shape[0]
    o2, a2, _ = build_coupler(4, device=None, x_range=(nx // 2, nx))
    assert a2.fields.E.shape[1] == nx - nx // 2 and "pml_min_x" not in a2.fields.psi_E and "pml_max_x" in a2.fields.psi_E


def test_array_container_reset_and_aset():
    from scenes import build_scene, seed_fields

    objects, arrays, cfg = build_scene(poles=1, detectors=("energy",))
    seed_fields(arrays)
    arrays.detector_states["energy"]["energy"][...] = 1
    eps = arrays.inv_permittivities.copy()
    r = arrays.reset()
    assert not r.fields.E.any() and not r.fields.dispersive_P_curr.any() and not r.detector_states["energy"]["energy"].any()
    assert np.array_equal(r.inv_permittivities, eps) and arrays.fields.E.any()
    r2 = r.aset("fields->E", np.ones_like(r.fields.E))
    assert r2.fields.E.all() and not r.fields.E.any()


def test_straddling_detector_states_shard_and_merge_round_trip():
    """a detector whose region crosses x-slab edges keeps, per rank, the state of its
    part of the region; ``merge_detector_states`` rebuilds the whole-region arrays (concatenation on x, sums
    for volume reductions, plane-count-weighted mean for the YZ mean of an energy slice set)."""
    import numpy as np

    import fdtdx_b200 as fx
    from fdtdx_b200.dist import merge_detector_states, shard_arrays, slab_bounds
    from fdtdx_b200.plan import clip_detector
    from scenes import build_scene

    nxs, world = 24, 3
    objects, arrays, cfg = build_scene(shape=(nxs, 6, 8), thickness=2, time=3e-15)
    wc = fx.WaveCharacter(wavelength=0.8e-6)
    dets = [
        fx.EnergyDetector(name="video", grid_slice_tuple=((0, nxs), (0, 6), (0, 8)), as_slices=True),
        fx.EnergyDetector(name="cube", grid_slice_tuple=((3, nxs - 2), (1, 5), (1, 7))),
        fx.FieldDetector(name="fmean", grid_slice_tuple=((2, nxs - 3), (1, 5), (1, 7)), reduce_volume=True, components=("Ex", "Hz")),
        fx.PoyntingFluxDetector(name="flux", grid_slice_tuple=((0, nxs), (0, 6), (5, 6)), direction="+"),
        fx.PhasorDetector(name="ph", grid_slice_tuple=((1, nxs - 1), (3, 4), (1, 7)), wave_characters=(wc,)),
        fx.EnergyDetector(name="fixed", grid_slice_tuple=((0, nxs), (0, 6), (0, 8)), as_slices=True, x_slice=13 * 50e-9, y_slice=2 * 50e-9, z_slice=3 * 50e-9),
        fx.FieldDetector(name="inside", grid_slice_tuple=((9, 15), (1, 5), (1, 7))),
    ]
    objects, arrays, _, cfg, _ = fx.place_objects(list(objects.object_list) + dets, cfg, inv_permittivities=arrays.inv_permittivities)
    rng = np.random.default_rng(0)
    for st in arrays.detector_states.values():
        for k, v in st.items():
            v[...] = rng.standard_normal(v.shape).astype(v.dtype)
    bounds = [slab_bounds(nxs, world, r) for r in range(world)]
    parts = [shard_arrays(arrays, objects, b, config=cfg).detector_states for b in bounds]
    # per-rank parts have the clipped shapes; the fixed-x plane of "fixed" lives on one rank only
    assert parts[1]["cube"]["energy"].shape[1] == 8 and parts[0]["cube"]["energy"].shape[1] == 5
    assert [clip_detector(objects["fixed"], *b, cfg)._slice_indices[0] for b in bounds] == [-1, objects["fixed"]._slice_indices[0] - 8, -1]
    assert "inside" in parts[1] and "inside" not in parts[0]
    # mean over x of the YZ plane: emulate what each rank's kernel writes (the mean over ITS planes)
    merged = merge_detector_states(objects, cfg, bounds, parts)
    for d, st in arrays.detector_states.items():
        for k, v in st.items():
            if d == "video" and k == "YZ Plane":
                continue  # parts hold the global value on rank 0 and zeros elsewhere: a weighted mean of those is not the identity
            np.testing.assert_allclose(merged[d][k], v, rtol=1e-6, atol=1e-7, err_msg=f"{d}/{k}")
    per_rank = [{"video": {"YZ Plane": np.full((1, 6, 8), float(r + 1), np.float32), **{k: p["video"][k] for k in ("XY Plane", "XZ Plane")}}} for r, p in enumerate(parts)]
    full = [{**p, **q} for p, q in zip(parts, per_rank)]
    yz = merge_detector_states(objects, cfg, bounds, full)["video"]["YZ Plane"]
    np.testing.assert_allclose(yz, 2.0, rtol=1e-6)  # (1 + 2 + 3) * 8 / 24
