"""Fixtures produced by the REFERENCE'S OWN SOURCE for the hot path (tests/golden/ref_steps.npz).

    python tests/golden/make_reference_golden.py          # needs /root/reference (this container only)

``oracle/refexec.py`` executes the reference's files (``fdtd/update.py``, ``core/physics/curl.py``,
``core/physics/metrics.py``, ``fdtd/misc.py``, ``objects/boundaries/*.py``, ``objects/detectors/*.py``,
``objects/sources/{tfsf,dipole,profile,source}.py`` ...) unchanged, with ``jax.numpy`` replaced by a NumPy-backed
stand-in (float32 / complex64 like JAX without x64) and this repo's host-mirror objects carrying the reference
classes' methods.  Each scene below is stepped with the reference's ``update_E`` / ``update_H`` /
``update_detector_states`` (and the reverse updates, the Bloch pad correction, the symmetry mirror); the final
fields, CPML psi, ADE polarisations and detector states are stored.  ``tests/test_reference_golden.py`` checks the
oracle (CPU) and the CUDA kernels (GPU) against them.  Scenes are rebuilt from their seeds by ``scene(name)``."""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

PATH = os.path.join(HERE, "ref_steps.npz")
ALL_DETS = ("field", "energy", "energy_slices", "energy_pos", "energy_reduce", "poynting", "poynting_full", "poynting_all", "phasor", "phasor_reduce", "field_reduce", "raw_field")

# name -> (builder, kwargs, forward steps, reverse steps)
SCENES = {
    "pml_iso": ("scene", dict(), 6, 0),
    "pml_kappa_nonuniform": ("scene", dict(kappa=True, nonuniform=True), 6, 0),
    "periodic_diag_sigma_mu": ("scene", dict(boundaries="periodic", eps_tier=3, sigma_E=True, mu_tier=3, sigma_H=True), 6, 0),
    "pec_pmc_walls": ("scene", dict(boundaries={"min_x": "pec", "max_x": "pmc", "min_y": "pml", "max_y": "pml", "min_z": "pmc", "max_z": "pec"}), 6, 0),
    "tensor_eps9_mu9": ("scene", dict(eps_tier=9, mu_tier=9, nonuniform=True), 5, 0),
    "tensor_eps9_sigma9": ("scene", dict(eps_tier=9, sigma_E=9), 5, 0),
    "ade_2poles_c4_sigma": ("scene", dict(poles=2, c4=True, sigma_E=True), 6, 0),
    "ade_1pole_diag_coeffs": ("scene", dict(poles=1, eps_tier=3, coeff_tier=3), 6, 0),
    "ade_in_tensor": ("scene", dict(poles=1, eps_tier=9), 5, 0),
    "detectors_all_nonuniform": ("scene", dict(detectors=ALL_DETS, nonuniform=True, source="plane_z"), 8, 0),
    "detectors_periodic": ("scene", dict(detectors=("field", "energy_slices", "poynting", "phasor"), boundaries="periodic", source="plane_z"), 8, 0),
    "source_plane_z_tilted": ("scene", dict(source="plane_z", nonuniform=True), 8, 0),
    "source_pulse": ("scene", dict(source="pulse"), 8, 0),
    "source_gated_switch": ("scene", dict(source="gated"), 10, 0),
    "source_table": ("scene", dict(source="table"), 8, 0),
    "source_plane_x_diag": ("scene", dict(source="plane_x", eps_tier=3, mu_tier=3), 8, 0),
    "source_plane_y_tensor": ("scene", dict(source="plane_y", eps_tier=9), 6, 0),
    "dipole": ("scene", dict(source="dipole"), 8, 0),
    "reverse_sigma_source": ("scene", dict(source="plane_z", sigma_E=True, sigma_H=True, mu_tier=1), 5, 5),
    "reverse_tensor": ("scene", dict(eps_tier=9, mu_tier=9, sigma_E=9), 4, 4),
    "bloch_xy": ("bloch", dict(shape=(10, 8, 16), types="BXY", k=(1e6, 1e6, 0.0), source=True, detectors=("energy", "poynting", "phasor")), 8, 0),
    "bloch_z_ragged": ("bloch", dict(shape=(6, 12, 9), types="BZ", k=(0.0, 0.0, 1.7e6), source=True, detectors=("energy", "poynting")), 8, 0),
    "symmetry_electric_xy": ("symmetry", dict(case="electric_xy_far_periodic"), 8, 0),
    "symmetry_magnetic_x": ("symmetry", dict(case="magnetic_x_far_periodic"), 8, 0),
}


def scene(name):
    """(objects, arrays, config) of a fixture scene, rebuilt from its seeds."""
    kind, kw, _, _ = SCENES[name]
    if kind == "scene":
        from scenes import build_scene, seed_fields

        objects, arrays, cfg = build_scene(**kw)
        seed_fields(arrays, seed=11)
    elif kind == "bloch":
        import test_bloch as tb

        kw = dict(kw)
        types = getattr(tb, kw.pop("types"))
        objects, arrays, cfg = tb.build(kw.pop("shape"), types, kw.pop("k"), time=5e-15, **kw)
        tb.seed_complex(arrays, seed=11)
    else:
        import test_symmetry as ts
        from scenes import seed_fields

        objects, arrays, cfg = ts.build((10, 8, 16), **ts.CASES[kw["case"]])
        seed_fields(arrays, seed=11)
    return objects, arrays, cfg


def flatten(arrays, prefix):
    out = {f"{prefix}/E": np.asarray(arrays.fields.E), f"{prefix}/H": np.asarray(arrays.fields.H)}
    for k, (a, b) in arrays.fields.psi_E.items():
        out[f"{prefix}/psi_E/{k}/0"], out[f"{prefix}/psi_E/{k}/1"] = np.asarray(a), np.asarray(b)
    for k, (a, b) in arrays.fields.psi_H.items():
        out[f"{prefix}/psi_H/{k}/0"], out[f"{prefix}/psi_H/{k}/1"] = np.asarray(a), np.asarray(b)
    if arrays.fields.dispersive_P_curr is not None:
        out[f"{prefix}/P_curr"], out[f"{prefix}/P_prev"] = np.asarray(arrays.fields.dispersive_P_curr), np.asarray(arrays.fields.dispersive_P_prev)
    for d, st in arrays.detector_states.items():
        for k, v in st.items():
            out[f"{prefix}/det/{d}/{k}"] = np.asarray(v)
    return out


def run_reference(ref, name):
    from oracle import refexec

    _, _, fwd, rev = SCENES[name]
    objects, arrays, cfg = scene(name)
    robj = ref.wrap_objects(objects, cfg)
    a = refexec.to_jarr(arrays)
    dets = len(objects.detectors) > 0
    for t in range(fwd):
        tt = ref.jnp.asarray(t, dtype=np.int32)
        H_prev = a.fields.H
        a = ref.update_E(tt, a, robj, cfg, True)
        a = ref.update_H(tt, a, robj, cfg, True)
        if dets:
            a = ref.update_detector_states(tt, a, robj, cfg, H_prev, False)
    out = flatten(a, f"{name}/fwd")
    if rev:
        for t in range(fwd - 1, fwd - 1 - rev, -1):  # backward.py:62-135 without interface replay / PML reset
            tt = ref.jnp.asarray(t, dtype=np.int32)
            a = ref.update_H_reverse(tt, a, robj, cfg)
            a = ref.update_E_reverse(tt, a, robj, cfg)
        out.update(flatten(a, f"{name}/rev"))
    return out


def run_oracle(name):
    from oracle import yee

    _, _, fwd, rev = SCENES[name]
    objects, arrays, cfg = scene(name)
    st = (0, arrays)
    dets = len(objects.detectors) > 0
    for _ in range(fwd):
        st = yee.forward(st, cfg, objects, record_detectors=dets)
    out = flatten(st[1], f"{name}/fwd")
    if rev:
        a = st[1]
        for t in range(fwd - 1, fwd - 1 - rev, -1):
            a = yee.update_H_reverse(t, a, objects, cfg)
            a = yee.update_E_reverse(t, a, objects, cfg)
        out.update(flatten(a, f"{name}/rev"))
    return out


def extra_reference(ref):
    """Pure functions pinned directly: interface gather / scatter (fdtd/misc.py:10-66), co-location stencil on a
    stretched grid (curl.py:86-224), energy / Poynting densities (metrics.py:15-117)."""
    from oracle import refexec
    from scenes import build_scene, seed_fields

    objects, arrays, cfg = build_scene(nonuniform=True, eps_tier=3, mu_tier=3)
    seed_fields(arrays, seed=5)
    robj = ref.wrap_objects(objects, cfg)
    a = refexec.to_jarr(arrays)
    out = {}
    vals = ref.collect_boundary_interfaces(a, robj.pml_objects)
    for k, v in vals.items():
        out[f"extra/interfaces/{k}"] = np.asarray(v)
    doubled = {k: 2.0 * v for k, v in vals.items()}
    b = ref.add_boundary_interfaces(a, doubled, robj.pml_objects)
    out["extra/interfaces_added/E"], out["extra/interfaces_added/H"] = np.asarray(b.fields.E), np.asarray(b.fields.H)
    Ei, Hi = ref.interpolate_fields(ref.pad_fields(a.fields.E, (False, True, False)), ref.pad_fields(a.fields.H, (False, True, False)), config=cfg)
    out["extra/interp/E"], out["extra/interp/H"] = np.asarray(Ei), np.asarray(Hi)
    out["extra/energy"] = np.asarray(ref.compute_energy(a.fields.E, a.fields.H, a.inv_permittivities, a.inv_permeabilities))
    out["extra/poynting"] = np.asarray(ref.compute_poynting_flux(a.fields.E, a.fields.H))
    return out


PML_CASES = [(nonuni, kappa) for nonuni in (False, True) for kappa in (False, True)]


def _pml_hosts(nonuni, kappa):
    """Fresh host PML objects (user fields only; sigma_end left to be derived) of a 14x12x18 scene, 5-cell slabs."""
    import fdtdx_b200 as fx
    from scenes import build_scene

    objects, _, cfg = build_scene(shape=(14, 12, 18), thickness=5, nonuniform=nonuni)
    hosts = []
    for pml in objects.pml_objects:
        h = fx.PerfectlyMatchedLayer(name=pml.name, grid_slice_tuple=pml.grid_slice_tuple, axis=pml.axis, direction=pml.direction)
        if kappa:
            h.kappa_end = 4.0
        hosts.append(h)
    return hosts, cfg


def pml_tables_reference(ref):
    """CPML a / b / 1/kappa tables (E and H side) from the reference's own place_on_grid body."""
    from oracle import refexec

    out = {}
    for nonuni, kappa in PML_CASES:
        hosts, cfg = _pml_hosts(nonuni, kappa)
        for h in hosts:
            for k, v in refexec.reference_pml_tables(ref, h, cfg).items():
                out[f"pml_tables/nonuniform={int(nonuni)},kappa={int(kappa)}/{h.name}/{k}"] = v
    return out


def pml_tables_host():
    """The same tables from this repo's host mirror (fdtdx_b200/boundaries.py) - what oracle AND kernels consume."""
    out = {}
    for nonuni, kappa in PML_CASES:
        hosts, cfg = _pml_hosts(nonuni, kappa)
        for h in hosts:
            h.place_on_grid(cfg)
            for k in ("pml_a_E", "pml_b_E", "inv_kappa_E", "pml_a_H", "pml_b_H", "inv_kappa_H"):
                out[f"pml_tables/nonuniform={int(nonuni)},kappa={int(kappa)}/{h.name}/{k}"] = np.asarray(getattr(h, k))
    return out


def extra_oracle():
    from oracle import yee
    from scenes import build_scene, seed_fields

    objects, arrays, cfg = build_scene(nonuniform=True, eps_tier=3, mu_tier=3)
    seed_fields(arrays, seed=5)
    out = {}
    for fs in ("E", "H"):
        for pml in objects.pml_objects:
            out[f"extra/interfaces/{pml.name}_{fs}"] = getattr(arrays.fields, fs)[(slice(None), *pml.interface_slice())]
    E, H = arrays.fields.E.copy(), arrays.fields.H.copy()
    for fs, arr in (("E", E), ("H", H)):
        for pml in objects.pml_objects:
            arr[(slice(None), *pml.interface_slice())] = 2.0 * out[f"extra/interfaces/{pml.name}_{fs}"]
    out["extra/interfaces_added/E"], out["extra/interfaces_added/H"] = E, H
    Ei, Hi = yee.interpolate_fields(yee.pad_fields(arrays.fields.E, (False, True, False)), yee.pad_fields(arrays.fields.H, (False, True, False)), config=cfg)
    out["extra/interp/E"], out["extra/interp/H"] = Ei, Hi
    out["extra/energy"] = yee.compute_energy(arrays.fields.E, arrays.fields.H, arrays.inv_permittivities, arrays.inv_permeabilities)
    out["extra/poynting"] = yee.compute_poynting_flux(arrays.fields.E, arrays.fields.H)
    return out


def main():
    from oracle import refexec

    ref = refexec.Reference()
    out = {}
    for name in SCENES:
        out.update(run_reference(ref, name))
        print(f"{name}: done", flush=True)
    out.update(extra_reference(ref))
    out.update(pml_tables_reference(ref))
    np.savez_compressed(PATH, **out)
    print(f"wrote {PATH}: {len(out)} arrays, {os.path.getsize(PATH) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
