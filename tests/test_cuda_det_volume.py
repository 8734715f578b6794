"""Row-marching detector kernels (csrc/det_volume.cuh) for large exact-interpolation regions.

They must (i) match the CPU oracle within BASELINE.json's detector bound (<= 1e-4) and (ii) reproduce
the generic one-thread-per-cell kernels: bit for bit wherever a value is written per cell or reduced
by the shared reduction kernels, and to float32 rounding (<= 2e-6) for the averaged energy slices,
whose sums are folded in a different - fixed - order."""

import numpy as np
import pytest
import torch

import fdtdx_b200 as fx
from oracle import yee
from scenes import build_scene, rel_l2

pytestmark = pytest.mark.gpu

CASES = {
    # full-volume videos + reductions, regions touching every face (zero halos of the co-location stencil)
    "pml_videos": dict(shape=(20, 18, 40), thickness=4, source="plane_z", time=8e-15,
                       detectors=("energy_slices", "energy", "energy_reduce", "field", "poynting_all", "phasor_reduce", "phasor")),
    # wrap halos on every axis, inverse detector present (ignored in the forward pass)
    "periodic": dict(shape=(16, 12, 36), boundaries="periodic", source="plane_z", time=6e-15, detectors=("energy_slices", "field", "energy_reduce", "inverse_energy")),
    # stretched grid (width-weighted backward averages), two z tiles (Nz > 128), region offset lo_z = 1
    "nonuniform_two_tiles": dict(shape=(10, 12, 136), thickness=3, source="plane_x", nonuniform=True, time=6e-15, detectors=("energy_slices", "field", "phasor_reduce", "poynting_all")),
    # diagonal eps / mu tiers in the energy density
    "diag_materials": dict(shape=(14, 16, 32), thickness=3, source="plane_z", eps_tier=3, mu_tier=3, time=6e-15, detectors=("energy_slices", "energy", "energy_reduce")),
}


def _run(kw, volume, monkeypatch):
    monkeypatch.setenv("FDTDX_B200_DET_VOLUME", "1" if volume else "0")
    objects, arrays, cfg = build_scene(**kw)
    t_end, out = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg)
    torch.cuda.synchronize()
    from fdtdx_b200.fdtd import get_plan

    plan = next(iter(objects.__dict__["_plan_cache"].values()))
    return objects, arrays, cfg, out, plan


@pytest.mark.parametrize("name", list(CASES))
def test_volume_kernels_match_oracle_and_generic_kernels(name, monkeypatch):
    kw = CASES[name]
    objects, arrays, cfg, out_v, plan_v = _run(kw, True, monkeypatch)
    _, _, _, out_g, plan_g = _run(kw, False, monkeypatch)
    ref = yee.checkpointed_fdtd(arrays, objects, cfg)[1]
    n_checked = 0
    for dname, st in ref.detector_states.items():
        for key, want in st.items():
            got_v = out_v.detector_states[dname][key].cpu().numpy()
            got_g = out_g.detector_states[dname][key].cpu().numpy()
            if np.abs(want).max() == 0:
                assert np.abs(got_v).max() == 0
                continue
            assert rel_l2(got_v, want) <= 1e-4, (dname, key, rel_l2(got_v, want))
            if "Plane" in key and dname != "energy_pos":
                assert rel_l2(got_v, got_g) <= 2e-6, (dname, key, rel_l2(got_v, got_g))
            else:
                assert np.array_equal(got_v, got_g), (dname, key, rel_l2(got_v, got_g))
            n_checked += 1
    assert n_checked >= 3
    assert torch.equal(out_v.fields.E, out_g.fields.E)
    # the two runs really took different detector paths
    assert plan_v.launch_count() != plan_g.launch_count()


def test_volume_path_in_the_reverse_pass(monkeypatch):
    """Inverse detectors are sampled during full_backward with H_prev = H before the reverse H update."""
    kw = dict(shape=(16, 14, 32), thickness=3, source="plane_z", time=5e-15, detectors=("energy_slices", "inverse_energy"), recorder=fx.Recorder(modules=[]))
    outs = []
    for vol in (True, False):
        monkeypatch.setenv("FDTDX_B200_DET_VOLUME", "1" if vol else "0")
        objects, arrays, cfg = build_scene(**kw)
        state = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg)
        state = fx.full_backward(state, objects, cfg, record_detectors=True, reset_fields=True)
        outs.append(state[1])
    st = yee.checkpointed_fdtd(arrays, objects, cfg)
    ref = yee.full_backward(st, objects, cfg)[1]
    for key, want in ref.detector_states["inverse_energy"].items():
        got = outs[0].detector_states["inverse_energy"][key].cpu().numpy()
        assert np.abs(want).max() > 0
        assert rel_l2(got, want) <= 1e-4
        assert rel_l2(got, outs[1].detector_states["inverse_energy"][key].cpu().numpy()) <= 2e-6
