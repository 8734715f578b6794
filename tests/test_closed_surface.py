"""Closed-surface and frequency-domain Poynting detectors (SURVEY.md section 8 a13;
objects/detectors/poynting_flux.py:199-284, 287-388, 391-540)."""

import numpy as np
import pytest

import fdtdx_b200 as fx
from oracle import yee
from scenes import build_scene, rel_l2

KW = dict(shape=(18, 16, 20), thickness=3, source="dipole", time=8e-15, detectors=("closed_flux", "closed_flux_in", "closed_phasor", "phasor_flux", "poynting"))


def test_oracle_closed_surface_flux_is_the_sum_of_its_faces():
    """Known answer: the box flux equals the signed sum of six single-plane PoyntingFluxDetectors on the
    box's boundary cell layers (the reference's own definition, metrics.py:120-160)."""
    objects, arrays, cfg = build_scene(**KW)
    box = next(d for d in objects.detectors if d.name == "closed_flux").grid_slice_tuple
    planes = []
    for a in range(3):
        for side, sgn in ((box[a][0], -1.0), (box[a][1] - 1, +1.0)):
            sl = list(box)
            sl[a] = (side, side + 1)
            planes.append((fx.PoyntingFluxDetector(name=f"f{a}{side}", grid_slice_tuple=tuple(sl), direction="+"), sgn))
    kw = dict(KW, detectors=("closed_flux",))
    objects2, arrays2, cfg2 = build_scene(**kw)
    objs = list(objects2.object_list) + [p for p, _ in planes]
    objects2, arrays2, _, cfg2, _ = fx.place_objects(objs, cfg2, inv_permittivities=arrays2.inv_permittivities)
    st = yee.checkpointed_fdtd(arrays2, objects2, cfg2)[1].detector_states
    total = sum(sgn * st[p.name]["poynting_flux"][:, 0] for p, sgn in planes)
    got = st["closed_flux"]["poynting_flux"][:, 0]
    assert np.abs(got).max() > 0
    np.testing.assert_allclose(got, total, rtol=2e-4, atol=2e-6 * np.abs(got).max())
    # a dipole inside a lossless box radiates outward: positive net flux once the wave has crossed the faces
    assert got[-1] > 0


@pytest.mark.gpu
def test_cuda_closed_surface_detectors_match_oracle():
    objects, arrays, cfg = build_scene(**KW)
    ref = yee.checkpointed_fdtd(arrays, objects, cfg)[1]
    t_end, out = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg)
    for name in ("closed_flux", "closed_flux_in", "poynting"):
        want = ref.detector_states[name]["poynting_flux"]
        got = out.detector_states[name]["poynting_flux"].cpu().numpy()
        assert np.abs(want).max() > 0
        assert rel_l2(got, want) <= 1e-4, (name, rel_l2(got, want))
    for key, want in ref.detector_states["closed_phasor"].items():
        got = out.detector_states["closed_phasor"][key].cpu().numpy()
        assert np.abs(want).max() > 0 and rel_l2(got, want) <= 1e-4, key
    assert rel_l2(out.detector_states["phasor_flux"]["phasor"].cpu().numpy(), ref.detector_states["phasor_flux"]["phasor"]) <= 1e-4
    # post-run integrals: same numbers from device states and from oracle states
    dets = {d.name: d for d in objects.detectors}
    nf_gpu = dets["closed_phasor"].compute_net_flux(out.detector_states["closed_phasor"])
    nf_ref = dets["closed_phasor"].compute_net_flux(ref.detector_states["closed_phasor"])
    assert rel_l2(nf_gpu, nf_ref) <= 1e-4 and nf_ref[0] > 0
    pf_gpu = dets["phasor_flux"].compute_poynting_flux(out.detector_states["phasor_flux"])
    pf_ref = dets["phasor_flux"].compute_poynting_flux(ref.detector_states["phasor_flux"])
    assert pf_gpu.shape == (2,) and rel_l2(pf_gpu, pf_ref) <= 1e-4


@pytest.mark.gpu
def test_closed_surface_flux_gradient_equals_the_gradient_of_its_six_faces():
    """reversible_fdtd gradient through ClosedSurfacePoyntingFluxDetector (det_adjoint_kernel, DET_CLOSED branch): the
    box flux is the signed sum of six single-plane fluxes (known answer above), so a loss on the box detector and the
    same loss on the six planes must give the same d loss / d inv_eps."""
    import torch

    rec = fx.Recorder(modules=[])
    kw = dict(KW, detectors=("closed_flux",), recorder=rec)
    objects, arrays, cfg = build_scene(**kw)
    box = next(d for d in objects.detectors if d.name == "closed_flux").grid_slice_tuple
    planes = []
    for a in range(3):
        for side, sgn in ((box[a][0], -1.0), (box[a][1] - 1, +1.0)):
            sl = list(box)
            sl[a] = (side, side + 1)
            planes.append((fx.PoyntingFluxDetector(name=f"f{a}{side}", grid_slice_tuple=tuple(sl), direction="+"), sgn))
    objs = list(objects.object_list) + [p for p, _ in planes]
    objects, arrays, _, cfg, _ = fx.place_objects(objs, cfg, inv_permittivities=arrays.inv_permittivities)
    T = cfg.time_steps_total

    def grad(terms):
        dev = arrays.to_torch("cuda")
        dev.inv_permittivities.requires_grad_(True)
        _, out = fx.run_fdtd(dev, objects, cfg)
        w = torch.linspace(0.5, 1.5, T, device="cuda")  # a cotangent that varies over the recorded steps
        loss = sum(sgn * (out.detector_states[n]["poynting_flux"][:, 0] * w).sum() for n, sgn in terms) * 1e18
        loss.backward()
        return dev.inv_permittivities.grad.cpu().numpy(), float(loss.detach())

    g_box, l_box = grad([("closed_flux", 1.0)])
    g_faces, l_faces = grad([(p.name, sgn) for p, sgn in planes])
    assert np.abs(g_faces).max() > 0 and abs(l_box - l_faces) <= 2e-4 * abs(l_faces)
    err = rel_l2(g_box, g_faces)
    print(f"closed-surface flux gradient vs six faces: rel-L2 {err:.2e}")
    assert err <= 1e-4
