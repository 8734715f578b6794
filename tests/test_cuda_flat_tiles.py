"""Flat tiles of the TMA-staged half-steps (csrc/yee_tma.cuh, TmaRt): on rows that are neither 64
nor 128 cells the CTA's threads are laid over (row, z quad) pairs.  Every scene is stepped three ways - flat tiles
(default), the 64- / 128-cell tile rows (FDTDX_B200_TMA_FLAT=0) and the register-marching kernels - and the three
results must be bit-identical; one case per geometry feature the flat mapping has to get right."""

import numpy as np
import pytest

from scenes import build_scene, rel_l2, seed_fields
from test_cuda_tma import _assert_identical, _np, _run
from oracle import yee

pytestmark = pytest.mark.gpu

PEC_PMC = {"min_x": "pml", "max_x": "pml", "min_y": "pec", "max_y": "pmc", "min_z": "pmc", "max_z": "pec"}
FLAT_SCENES = {
    # 18 quads per row: 14 rows per CTA, the last warp partly past the tile; Ny not a multiple of the tile rows
    "one_tile_72": (dict(shape=(11, 31, 72), thickness=4, source="plane_z", detectors=("energy_slices", "phasor")), 6),
    # rows of 5 quads: 51 rows per CTA, a warp spans 6-7 rows
    "short_rows_20": (dict(shape=(7, 60, 20), thickness=3, eps_tier=3, sigma_E=True), 5),
    # two equal flat tiles of 17 quads instead of 128 + 8 cells; z-slab coefficients per tile
    "two_tiles_136": (dict(shape=(8, 19, 136), thickness=3, kappa=True, nonuniform=True), 5),
    # three flat tiles, CPML thinner than a quad, diagonal mu with conductivity
    "three_tiles_260": (dict(shape=(5, 9, 260), thickness=2, mu_tier=3, sigma_H=True), 5),
    # PEC / PMC walls on the y and z faces, dispersive cells
    "walls_ade_76": (dict(shape=(8, 13, 76), thickness=3, boundaries=PEC_PMC, poles=1), 5),
}


@pytest.mark.parametrize("xchunk", [1, 3, 0])
@pytest.mark.parametrize("name", list(FLAT_SCENES))
def test_flat_tiles_equal_fixed_tiles_marching_and_oracle(name, xchunk, monkeypatch):
    kw, steps = FLAT_SCENES[name]
    objects, arrays, cfg = build_scene(**kw)
    seed_fields(arrays, seed=11)
    rec = bool(kw.get("detectors"))
    marching = _run(objects, arrays, cfg, steps, tma=0, xchunk=0, record_detectors=rec)
    flat = _run(objects, arrays, cfg, steps, tma=1, xchunk=xchunk, record_detectors=rec)
    monkeypatch.setenv("FDTDX_B200_TMA_FLAT", "0")
    fixed = _run(objects, arrays, cfg, steps, tma=1, xchunk=xchunk, record_detectors=rec)
    _assert_identical(marching, flat)
    _assert_identical(fixed, flat)
    for dname, st in flat.detector_states.items():
        for key, v in st.items():
            assert np.array_equal(_np(v), _np(marching.detector_states[dname][key])), (dname, key)
    if xchunk == 0 and not rec:
        st = (0, arrays)
        for _ in range(steps):
            st = yee.forward(st, cfg, objects, None, False, False, True)
        assert rel_l2(_np(flat.fields.E), st[1].fields.E) <= 1e-5
        assert rel_l2(_np(flat.fields.H), st[1].fields.H) <= 1e-5
