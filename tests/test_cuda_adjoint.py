"""reversible_fdtd gradients: the fused adjoint kernels vs the gradient of the
differentiable CPU oracle (torch autograd through ``oracle/yee_torch.py`` in float64 - the
"checkpointed" reference the reference's own test compares against, tests/simulation/fdtd/test_fdtd.py:440-498).
Tolerance: BASELINE.json north_star, time-reversal gradients <= 1e-4 relative L2.

Two oracles: (i) without CPML the reversible method is exact, so the product must match plain
autograd of the forward run (the reference's own reversible-vs-checkpointed cross-check uses
periodic scenes); (ii) with CPML the reference's method is approximate by design (psi frozen,
slab fields not reconstructible, Poynting/energy detectors linearised at reconstructed fields -
SURVEY Appendix C.1), so the product is compared with ``yee_torch.reversible_gradient``, the
restatement of ``fdtd_bwd`` itself, on the cells outside the slabs (inside them rounding noise is
amplified by the reverse pass and neither side is meaningful)."""

import numpy as np
import pytest
import torch

import fdtdx_b200 as fx
from oracle import yee_torch
from scenes import build_scene, rel_l2

pytestmark = pytest.mark.gpu


def _loss(det, E=None):
    loss = 0.0
    for name, st in det.items():
        for k, v in st.items():
            if v.is_complex():
                loss = loss + (v.real**2 + v.imag**2).sum() * 3.0 + v.real.sum() * 1e-3
            elif "poynting" in k:
                loss = loss + v.sum() * 1e19
            elif "energy" in k or "Plane" in k:
                loss = loss + v.sum() * 1e21 if v.numel() < 1000 else loss + (v * v).sum() * 1e3 + v.sum()
            else:
                loss = loss + (v * v).sum() + v.sum() * 1e-2
    if E is not None:
        loss = loss + (E * E).sum() * 0.1
    return loss


CASES = {
    "ragged_nz21_pml_poynting": (dict(source="plane_z", detectors=("poynting", "phasor", "field_reduce"), time=4e-15, shape=(16, 14, 21), thickness=4), 5, False),
    "periodic_field_phasor": (dict(boundaries="periodic", source="plane_z", detectors=("field", "phasor"), time=3e-15), None, True),
    "pml_poynting_phasor": (dict(source="plane_z", detectors=("poynting", "phasor", "field_reduce"), time=4e-15, shape=(16, 14, 20), thickness=4), 5, False),
    "pml_energy_sigma_diag": (dict(source="plane_z", detectors=("energy_reduce", "poynting_full", "energy_slices"), eps_tier=3, sigma_E=True, sigma_H=True, time=3e-15, shape=(16, 14, 20), thickness=4), 5, False),
    "nonuniform_pec_dipole": (dict(source="dipole", detectors=("poynting_all", "phasor_reduce", "raw_field"), nonuniform=True, time=3e-15, shape=(14, 12, 16),
                                   boundaries={"min_x": "pec", "max_x": "pmc", "min_y": "periodic", "max_y": "periodic", "min_z": "pec", "max_z": "pmc"}), None, True),
    "mu_array_kappa": (dict(source="plane_z", detectors=("poynting", "field"), mu_tier=3, kappa=True, time=3e-15, shape=(16, 14, 20), thickness=4), 5, False),
}


@pytest.mark.parametrize("name", list(CASES))
def test_reversible_gradient_matches_oracle(name):
    from oracle import yee

    kw, margin, final_E = CASES[name]
    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build_scene(recorder=rec, **kw)
    T = cfg.time_steps_total
    mu_np = arrays.inv_permeabilities
    has_mu = isinstance(mu_np, np.ndarray)
    loss_fn = lambda E, H, det: _loss(det, E if final_E else None)
    if margin is None:
        # exact case: float64 autograd through the restated forward run
        ie = torch.tensor(arrays.inv_permittivities.astype(np.float64), requires_grad=True)
        im = torch.tensor(mu_np.astype(np.float64), requires_grad=True) if has_mu else None
        E, H, det = yee_torch.run_forward(arrays.reset(), objects, cfg, T, inv_eps=ie, inv_mu=im, dtype=torch.float64)
        loss_ref = loss_fn(E, H, det)
        loss_ref.backward()
        g_ref, gm_ref = ie.grad.numpy(), (im.grad.numpy() if has_mu else None)
    else:
        st = yee.checkpointed_fdtd(arrays, objects, cfg)
        S = yee_torch.Stepper(st[1], objects, cfg)
        E, H, _, _, det = S.initial(st[1])
        loss_ref = loss_fn(E, H, det)
        g_ref, gm_ref = yee_torch.reversible_gradient(st[1], objects, cfg, loss_fn)
        g_ref, gm_ref = g_ref.numpy(), (gm_ref.numpy() if gm_ref is not None else None)
    # product: reversible_fdtd wired into torch autograd
    dev = arrays.to_torch("cuda")
    dev.inv_permittivities.requires_grad_(True)
    if has_mu:
        dev.inv_permeabilities.requires_grad_(True)
    t_end, out = fx.run_fdtd(dev, objects, cfg)
    assert t_end == T
    loss = _loss(out.detector_states, out.fields.E if final_E else None)
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    E_final = out.fields.E.detach().clone()
    loss.backward()
    assert torch.equal(out.fields.E.detach(), E_final), "backward must not disturb the returned fields"
    g = dev.inv_permittivities.grad.cpu().numpy()
    sl = (slice(None),) * 4 if margin is None else (slice(None), *(slice(margin, -margin),) * 3)
    assert np.abs(g_ref[sl]).max() > 0
    err = rel_l2(g[sl], g_ref[sl])
    assert err <= 1e-4, f"d loss / d inv_eps rel-L2 {err}"
    if has_mu:
        gm = dev.inv_permeabilities.grad.cpu().numpy()
        errm = rel_l2(gm[sl], gm_ref[sl])
        assert errm <= 1e-4, f"d loss / d inv_mu rel-L2 {errm}"


def test_sliced_checkpoints_match_single_slice():
    """num_checkpoints_reversible (fdtd.py:106-166, tests/simulation/fdtd/test_fdtd.py:618-700): exact
    field checkpoints at the slice boundaries leave the (periodic, exactly reversible) gradient unchanged."""
    grads = []
    for nck in (0, 3):
        rec = fx.Recorder(modules=[])
        objects, arrays, cfg = build_scene(recorder=rec, boundaries="periodic", source="plane_z", detectors=("field", "phasor"), time=3e-15, sigma_E=True)
        cfg = cfg.aset("gradient_config", fx.GradientConfig(method="reversible", recorder=rec, num_checkpoints_reversible=nck))
        dev = arrays.to_torch("cuda")
        dev.inv_permittivities.requires_grad_(True)
        _, out = fx.run_fdtd(dev, objects, cfg)
        _loss(out.detector_states, out.fields.E).backward()
        grads.append(dev.inv_permittivities.grad.cpu().numpy())
    assert rel_l2(grads[1], grads[0]) <= 1e-4


def test_reversible_needs_recorder_and_rejects_dispersion():
    objects, arrays, cfg = build_scene(poles=1, recorder=fx.Recorder(modules=[]))
    dev = arrays.to_torch("cuda")
    dev.inv_permittivities.requires_grad_(True)
    with pytest.raises(NotImplementedError):
        fx.reversible_fdtd(dev, objects, cfg)


CKPT_CASES = {
    # ragged grid (Nz % 4 != 0): interleaved marching kernels + scalar adjoint kernels
    "ragged_nz21_pml": dict(source="plane_z", detectors=("poynting", "phasor"), time=3e-15, shape=(16, 14, 21), thickness=4, eps_tier=3),
    "pml_poynting_phasor": dict(source="plane_z", detectors=("poynting", "phasor", "field_reduce"), time=4e-15, shape=(16, 14, 20), thickness=4),
    "pml_energy_sigma_diag_mu": dict(source="plane_z", detectors=("energy_reduce", "poynting_full"), eps_tier=3, sigma_E=True, mu_tier=3, time=3e-15, shape=(16, 14, 20), thickness=4),
    "nonuniform_kappa_dipole": dict(source="dipole", detectors=("poynting_all", "phasor_reduce"), nonuniform=True, kappa=True, time=3e-15, shape=(14, 12, 16), thickness=3),
}


@pytest.mark.parametrize("name", list(CKPT_CASES))
@pytest.mark.parametrize("nck", [1, 4])
def test_checkpointed_gradient_is_the_gradient_of_the_forward_run(name, nck):
    """GradientConfig(method="checkpointed") (fdtd.py:482-493): stored / recomputed states + the fused
    adjoint kernels at those states.  Unlike the reversible method this is exact with CPML, so it is
    compared with float64 autograd through the restated forward run on the WHOLE grid, slabs included
    (what the reference's reversible-vs-checkpointed test uses as its ground truth,
    tests/simulation/fdtd/test_fdtd.py:440-498)."""
    kw = CKPT_CASES[name]
    objects, arrays, cfg = build_scene(**kw)
    cfg = cfg.aset("gradient_config", fx.GradientConfig(method="checkpointed", num_checkpoints=nck))
    T = cfg.time_steps_total
    mu_np = arrays.inv_permeabilities
    has_mu = isinstance(mu_np, np.ndarray)
    ie = torch.tensor(arrays.inv_permittivities.astype(np.float64), requires_grad=True)
    im = torch.tensor(mu_np.astype(np.float64), requires_grad=True) if has_mu else None
    E, H, det = yee_torch.run_forward(arrays.reset(), objects, cfg, T, inv_eps=ie, inv_mu=im, dtype=torch.float64)
    loss_ref = _loss(det, E)
    loss_ref.backward()
    dev = arrays.to_torch("cuda")
    dev.inv_permittivities.requires_grad_(True)
    if has_mu:
        dev.inv_permeabilities.requires_grad_(True)
    t_end, out = fx.run_fdtd(dev, objects, cfg)
    assert t_end == T
    loss = _loss(out.detector_states, out.fields.E)
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    loss.backward()
    g, g_ref = dev.inv_permittivities.grad.cpu().numpy(), ie.grad.numpy()
    assert np.abs(g_ref).max() > 0
    assert rel_l2(g, g_ref) <= 1e-4, f"d loss / d inv_eps rel-L2 {rel_l2(g, g_ref)}"
    if has_mu:
        gm, gm_ref = dev.inv_permeabilities.grad.cpu().numpy(), im.grad.numpy()
        assert rel_l2(gm, gm_ref) <= 1e-4, f"d loss / d inv_mu rel-L2 {rel_l2(gm, gm_ref)}"



def _ckpt_setup(**kw):
    objects, arrays, cfg = build_scene(**kw)
    cfg = cfg.aset("gradient_config", fx.GradientConfig(method="checkpointed", num_checkpoints=2))
    mu_np = arrays.inv_permeabilities
    has_mu = isinstance(mu_np, np.ndarray)
    ie = torch.tensor(arrays.inv_permittivities.astype(np.float64), requires_grad=True)
    im = torch.tensor(mu_np.astype(np.float64), requires_grad=True) if has_mu else None
    E, H, det = yee_torch.run_forward(arrays.reset(), objects, cfg, cfg.time_steps_total, inv_eps=ie, inv_mu=im, dtype=torch.float64)
    dev = arrays.to_torch("cuda")
    dev.inv_permittivities.requires_grad_(True)
    if has_mu:
        dev.inv_permeabilities.requires_grad_(True)
    _, out = fx.run_fdtd(dev, objects, cfg)
    return det, out, ie, im, dev


@pytest.mark.parametrize("key", ["XY Plane", "XZ Plane", "YZ Plane"])
def test_loss_on_a_single_slice_plane(key):
    """Partial detector cotangents: a loss that uses one leaf of a multi-leaf detector state leaves the
    other leaves' cotangents unmaterialised (None).  The detector must still contribute, and the
    missing leaves must read as zero (they used to be dereferenced / the detector used to be skipped)."""
    det, out, ie, im, dev = _ckpt_setup(source="plane_z", detectors=("energy_slices", "poynting"), time=3e-15, shape=(16, 14, 20), thickness=4)
    f = lambda d: (d["energy_slices"][key] ** 2).sum() * 1e3 + d["energy_slices"][key].sum()
    f(det).backward()
    f(out.detector_states).backward()
    g, g_ref = dev.inv_permittivities.grad.cpu().numpy(), ie.grad.numpy()
    assert np.abs(g_ref).max() > 0
    assert rel_l2(g, g_ref) <= 1e-4


def test_conjugated_phasor_cotangent():
    """A loss built from ``phasor.conj()`` hands the backward pass a gradient with torch's lazy conj bit;
    it must be materialised before the kernels read it by pointer."""
    det, out, ie, im, dev = _ckpt_setup(source="plane_z", detectors=("phasor",), time=3e-15, shape=(16, 14, 20), thickness=4)
    w = torch.tensor(np.random.default_rng(3).standard_normal(det["phasor"]["phasor"].shape) + 1j * np.random.default_rng(4).standard_normal(det["phasor"]["phasor"].shape))
    f = lambda p, ww: (p.conj() * ww).imag.sum() + (p.conj() * p.conj()).real.sum() * 2.0
    f(det["phasor"]["phasor"], w).backward()
    f(out.detector_states["phasor"]["phasor"], w.to("cuda").to(torch.complex64)).backward()
    g, g_ref = dev.inv_permittivities.grad.cpu().numpy(), ie.grad.numpy()
    assert np.abs(g_ref).max() > 0
    assert rel_l2(g, g_ref) <= 1e-4


def test_energy_detector_gradient_wrt_inv_mu():
    """d(energy)/d(inv_mu) = -0.5 H^2 / inv_mu^2 joins grad_inv_mu when inv_permeabilities is an array
    (metrics.py:55-67 differentiated w.r.t. both materials)."""
    det, out, ie, im, dev = _ckpt_setup(source="plane_z", detectors=("energy_reduce", "energy_slices"), mu_tier=3, time=3e-15, shape=(16, 14, 20), thickness=4)
    f = lambda d: d["energy_reduce"]["energy"].sum() * 1e21 + (d["energy_slices"]["XZ Plane"] ** 2).sum() * 1e3
    f(det).backward()
    f(out.detector_states).backward()
    assert rel_l2(dev.inv_permittivities.grad.cpu().numpy(), ie.grad.numpy()) <= 1e-4
    gm, gm_ref = dev.inv_permeabilities.grad.cpu().numpy(), im.grad.numpy()
    assert np.abs(gm_ref).max() > 0
    assert rel_l2(gm, gm_ref) <= 1e-4


ADE_CASES = {
    "lorentz_1pole_iso": dict(poles=1, source="plane_z", detectors=("poynting", "phasor"), time=3e-15, shape=(16, 14, 20), thickness=4),
    "2poles_c4_sigma_diag": dict(poles=2, c4=True, coeff_tier=3, eps_tier=3, sigma_E=True, source="plane_z", detectors=("poynting", "field_reduce", "energy_reduce"), time=3e-15, shape=(16, 14, 20), thickness=4),
    "ragged_periodic_1pole_c4": dict(poles=1, c4=True, boundaries="periodic", source="plane_z", detectors=("field", "phasor"), time=3e-15, shape=(12, 10, 18)),
}


@pytest.mark.parametrize("name", list(ADE_CASES))
def test_checkpointed_gradient_through_dispersive_media(name):
    """ADE adjoint (the reference pins it in
    tests/simulation/fdtd/test_time_reversal.py:955-1044 with finite differences on c1, c2, c3): the
    checkpointed gradient w.r.t. inv_eps AND the recurrence coefficients c1..c4 equals float64 autograd
    through the restated dispersive forward run."""
    kw = ADE_CASES[name]
    objects, arrays, cfg = build_scene(**kw)
    cfg = cfg.aset("gradient_config", fx.GradientConfig(method="checkpointed", num_checkpoints=3))
    T = cfg.time_steps_total
    names = ["c1", "c2", "c3"] + (["c4"] if arrays.dispersive_c4 is not None else [])
    ie = torch.tensor(arrays.inv_permittivities.astype(np.float64), requires_grad=True)
    co = {k: torch.tensor(getattr(arrays, "dispersive_" + k).astype(np.float64), requires_grad=True) for k in names}
    E, H, det = yee_torch.run_forward(arrays.reset(), objects, cfg, T, inv_eps=ie, dtype=torch.float64, coeffs=co)
    loss_ref = _loss(det, E)
    loss_ref.backward()
    dev = arrays.to_torch("cuda")
    dev.inv_permittivities.requires_grad_(True)
    for k in names:
        getattr(dev, "dispersive_" + k).requires_grad_(True)
    t_end, out = fx.run_fdtd(dev, objects, cfg)
    assert t_end == T
    loss = _loss(out.detector_states, out.fields.E)
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    loss.backward()
    g, g_ref = dev.inv_permittivities.grad.cpu().numpy(), ie.grad.numpy()
    assert np.abs(g_ref).max() > 0
    assert rel_l2(g, g_ref) <= 1e-4, f"d loss / d inv_eps rel-L2 {rel_l2(g, g_ref)}"
    for k in names:
        gk, gk_ref = getattr(dev, "dispersive_" + k).grad.cpu().numpy(), co[k].grad.numpy()
        assert np.abs(gk_ref).max() > 0
        assert rel_l2(gk, gk_ref) <= 1e-4, f"d loss / d {k} rel-L2 {rel_l2(gk, gk_ref)}"


FUSED_CASES = {
    # two z tiles (128 + 8 cells), three y tiles, several x chunks; CPML on every face
    "pml_two_ztiles": dict(source="plane_z", detectors=("poynting", "phasor", "field_reduce"), time=3e-15, shape=(24, 19, 136), thickness=4),
    "pml_kappa_diag_mu": dict(source="plane_z", detectors=("poynting", "field"), eps_tier=3, mu_tier=3, kappa=True, time=3e-15, shape=(16, 14, 20), thickness=4),
    "periodic": dict(boundaries="periodic", source="plane_z", detectors=("field", "phasor"), time=3e-15),
    "nonuniform_mixed_walls": dict(source="dipole", detectors=("poynting_all", "phasor_reduce", "raw_field"), nonuniform=True, time=3e-15, shape=(14, 12, 16),
                                   boundaries={"min_x": "pec", "max_x": "pmc", "min_y": "periodic", "max_y": "periodic", "min_z": "pec", "max_z": "pmc"}),
    "pml_energy_slices": dict(source="plane_z", detectors=("energy_reduce", "poynting_full", "energy_slices"), time=3e-15, shape=(16, 14, 20), thickness=4),
}


def _grad_with_env(kw, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rec = fx.Recorder(modules=[])
    objects, arrays, cfg = build_scene(recorder=rec, **kw)
    dev = arrays.to_torch("cuda")
    dev.inv_permittivities.requires_grad_(True)
    has_mu = isinstance(arrays.inv_permeabilities, np.ndarray)
    if has_mu:
        dev.inv_permeabilities.requires_grad_(True)
    _, out = fx.run_fdtd(dev, objects, cfg)
    _loss(out.detector_states, out.fields.E).backward()
    torch.cuda.synchronize()
    return dev.inv_permittivities.grad.cpu().numpy(), (dev.inv_permeabilities.grad.cpu().numpy() if has_mu else None)


@pytest.mark.parametrize("name", list(FUSED_CASES))
def test_fused_adjoint_kernel_matches_two_kernel_form(name, monkeypatch):
    """adj_fused4_kernel (one x-marching pass, derivative cotangents re-evaluated at the neighbours) follows the
    arithmetic and summation order of adj_local4 + adj_gather4: same gradients up to the order of the detector
    kernels' atomic scatters (<= 1e-6)."""
    kw = FUSED_CASES[name]
    g0, m0 = _grad_with_env(kw, {"FDTDX_B200_ADJ_FUSED": "0", "FDTDX_B200_ADJ_INTERLEAVE": "1"}, monkeypatch)
    g1, m1 = _grad_with_env(kw, {"FDTDX_B200_ADJ_FUSED": "1", "FDTDX_B200_ADJ_INTERLEAVE": "1", "FDTDX_B200_ADJ_XC": "4"}, monkeypatch)
    assert np.abs(g0).max() > 0
    err = rel_l2(g1, g0)
    print(f"[{name}] fused vs two-kernel adjoint: d/d inv_eps rel-L2 {err:.2e}, max |delta| {np.abs(g1 - g0).max():.2e}")
    assert err <= 1e-6
    if m0 is not None:
        assert rel_l2(m1, m0) <= 1e-6


INTERLEAVE_CASES = {
    # nonlinear detectors (Poynting, energy) strictly inside the CPML slabs: every step runs interleaved
    "pml_interior_poynting_energy": dict(source="plane_z", detectors=("poynting_interior", "energy_interior", "phasor"), time=3e-15, shape=(24, 19, 136), thickness=4),
    "pml_diag_mu_interior": dict(source="plane_z", detectors=("poynting_interior", "field"), eps_tier=3, mu_tier=3, time=3e-15, shape=(18, 18, 24), thickness=4),
    # a Poynting plane that crosses the slabs: the steps it is on keep the reverse-then-recompute order
    "pml_poynting_across_slabs": FUSED_CASES["pml_two_ztiles"],
    "periodic": FUSED_CASES["periodic"],
}


@pytest.mark.parametrize("name", list(INTERLEAVE_CASES))
def test_interleaved_backward_matches_reverse_then_recompute(name, monkeypatch):
    """The interleaved backward iteration (reverse H -> transposes of the detectors and the H half-step -> reverse E ->
    transpose of the E half-step) uses the reconstructed E(t+1), H(t+1) where the reverse-then-recompute order re-runs
    the forward step: same gradient outside the CPML slabs up to the rounding of one reversed step."""
    kw = INTERLEAVE_CASES[name]
    g0, _ = _grad_with_env(kw, {"FDTDX_B200_ADJ_FUSED": "0", "FDTDX_B200_ADJ_INTERLEAVE": "0"}, monkeypatch)
    g1, _ = _grad_with_env(kw, {"FDTDX_B200_ADJ_FUSED": "1", "FDTDX_B200_ADJ_INTERLEAVE": "1"}, monkeypatch)
    th = kw.get("thickness", 0) + 1 if kw.get("boundaries") != "periodic" else 0
    sl = (slice(None),) * 4 if th == 0 else (slice(None), *(slice(th, -th),) * 3)
    err = rel_l2(g1[sl], g0[sl])
    print(f"[{name}] interleaved vs reverse-then-recompute: rel-L2 {err:.2e} outside the slabs")
    assert err <= 2e-5
