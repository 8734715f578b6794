"""Small seeded scenes shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""

from __future__ import annotations

import numpy as np

import fdtdx_b200 as fx

F = np.float32


def make_config(shape, spacing=50e-9, time=20e-15, nonuniform=False, recorder=None, seed=0):
    if nonuniform:
        rng = np.random.default_rng(seed + 17)
        edges = []
        for n in shape:
            w = spacing * (1.0 + 0.6 * rng.random(n))
            edges.append(np.concatenate([[0.0], np.cumsum(w)]))
        grid = fx.RectilinearGrid(*edges)
    else:
        grid = fx.UniformGrid(spacing=spacing)
    gc = None if recorder is None else fx.GradientConfig(method="reversible", recorder=recorder)
    return fx.SimulationConfig(time=time, grid=grid, gradient_config=gc)


def seed_fields(arrays, seed=0, amp=1e-3, psi=True, zero_in_pml=None):
    """E,H ~ amp * N(0,1) (mirrors ``_seed_fields`` of the reference's test_time_reversal.py)."""
    rng = np.random.default_rng(seed)
    arrays.fields.E[...] = (amp * rng.standard_normal(arrays.fields.E.shape)).astype(F)
    arrays.fields.H[...] = (amp * rng.standard_normal(arrays.fields.H.shape)).astype(F)
    if psi:
        for d in (arrays.fields.psi_E, arrays.fields.psi_H):
            for k, (a, b) in d.items():
                a[...] = (0.1 * amp * rng.standard_normal(a.shape)).astype(F)
                b[...] = (0.1 * amp * rng.standard_normal(b.shape)).astype(F)
    if arrays.fields.dispersive_P_curr is not None:
        arrays.fields.dispersive_P_curr[...] = (amp * rng.standard_normal(arrays.fields.dispersive_P_curr.shape)).astype(F)
        arrays.fields.dispersive_P_prev[...] = (amp * rng.standard_normal(arrays.fields.dispersive_P_prev.shape)).astype(F)
    return arrays


def build_scene(
    shape=(12, 10, 16),
    boundaries="pml",
    thickness=3,
    eps_tier=1,
    sigma_E=False,
    sigma_H=False,
    mu_tier=0,
    nonuniform=False,
    source=None,          # None | "plane_x" | "plane_y" | "plane_z" | "dipole" | "pulse" | "gated" | "table"
    detectors=(),         # names from: field, field_reduce, energy, energy_slices, energy_pos, energy_reduce, poynting, poynting_interior, energy_interior, poynting_full, poynting_all, phasor, phasor_reduce, raw_field, inverse_energy
    recorder=None,
    poles=0,
    c4=False,
    coeff_tier=1,
    time=None,
    seed=0,
    kappa=False,
):
    rng = np.random.default_rng(seed + 1)
    spacing = 50e-9
    cfg = make_config(shape, spacing=spacing, time=time or 12e-15, nonuniform=nonuniform, recorder=recorder, seed=seed)
    nx, ny, nz = shape
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=((0, nx), (0, ny), (0, nz)))
    bl = fx.boundary_objects_from_config(shape, cfg, boundaries, thickness=thickness)
    if kappa:
        for b in bl:
            if isinstance(b, fx.PerfectlyMatchedLayer):
                b.kappa_end = 4.0
                b.place_on_grid(cfg)
    inv_eps = (1.0 / (1.0 + 3.0 * rng.random((eps_tier, *shape)))).astype(F)
    if eps_tier == 9:
        # symmetric positive definite tensor per cell
        A = rng.standard_normal((3, 3, *shape)) * 0.2
        eps = np.einsum("ij...,kj...->ik...", A, A) + np.eye(3)[:, :, None, None, None] * (1.5 + rng.random(shape))
        inv_eps = np.linalg.inv(eps.transpose(2, 3, 4, 0, 1)).transpose(3, 4, 0, 1, 2).reshape(9, *shape).astype(F)
    inv_mu = 1.0
    if mu_tier:
        inv_mu = (1.0 / (1.0 + 0.5 * rng.random((mu_tier, *shape)))).astype(F)
        if mu_tier == 9:
            A = rng.standard_normal((3, 3, *shape)) * 0.1
            mu = np.einsum("ij...,kj...->ik...", A, A) + np.eye(3)[:, :, None, None, None] * (1.0 + 0.3 * rng.random(shape))
            inv_mu = np.linalg.inv(mu.transpose(2, 3, 4, 0, 1)).transpose(3, 4, 0, 1, 2).reshape(9, *shape).astype(F)
    sE = sH = None
    if sigma_E:
        t = sigma_E if isinstance(sigma_E, int) and sigma_E > 1 else (3 if eps_tier == 3 else 1)
        sE = (2e-4 * rng.random((t, *shape))).astype(F)
        if t == 9:
            sE = np.zeros((9, *shape), F)
            for d in (0, 4, 8):
                sE[d] = (2e-4 * rng.random(shape)).astype(F)
            sE[1] = sE[3] = (5e-5 * rng.random(shape)).astype(F)
    if sigma_H:
        t = 3 if mu_tier == 3 else 1
        sH = (30.0 * rng.random((t, *shape))).astype(F)
    objs = [vol, *bl]
    wc = fx.WaveCharacter(wavelength=0.8e-6)
    off = thickness + 1 if boundaries != "periodic" else 2
    if source in ("plane_x", "plane_y", "plane_z", "pulse", "gated", "table"):
        axis = {"plane_x": 0, "plane_y": 1, "plane_z": 2}.get(source, 2)
        sl = [(0, nx), (0, ny), (0, nz)]
        sl[axis] = (off, off + 1)
        prof, sw = None, None
        if source == "pulse":
            prof = fx.GaussianPulseProfile(spectral_width=fx.WaveCharacter(wavelength=4e-6), center_wave=wc)
        if source == "gated":
            sw = fx.OnOffSwitch(start_time=1e-15, end_time=6e-15)
        if source == "table":
            T = cfg.time_steps_total
            sig = np.sin(np.arange(T + 2) * 0.21).astype(F) * np.hanning(T + 2).astype(F)
            prof = fx.CustomTimeSignalProfile(signal=sig, time_step_duration=cfg.time_step_duration)
        pol = (1.0, 0.0, 0.0) if axis != 0 else (0.0, 1.0, 0.0)
        face = tuple(b - a for a, b in sl)
        amp = fx.gaussian_amplitude_profile(face, axis, radius_cells=0.45 * min(n for i, n in enumerate(face) if i != axis))
        src = fx.make_plane_source(
            "source", tuple(sl), cfg, inv_eps, inv_mu, direction="+" if axis != 1 else "-", wave_character=wc,
            temporal_profile=prof, fixed_E_polarization_vector=pol, amplitude_profile=amp,
            elevation_angle=10.0 if source == "plane_z" else 0.0, switch=sw,
        )
        objs.append(src)
    elif source == "dipole":
        c = (nx // 2, ny // 2, nz // 2)
        d = fx.PointDipoleSource(name="dipole", grid_slice_tuple=tuple((v, v + 1) for v in c), wave_character=wc, polarization=2, amplitude=1.0)
        objs.append(d)
    full = ((0, nx), (0, ny), (0, nz))
    inner = ((1, nx - 1), (2, ny - 2), (1, nz - 2))
    plane = ((0, nx), (0, ny), (nz - off - 2, nz - off - 1))
    xplane = ((nx - off - 2, nx - off - 1), (0, ny), (0, nz))
    mk = {
        "field": lambda: fx.FieldDetector(name="field", grid_slice_tuple=inner, switch=fx.OnOffSwitch(interval=2)),
        "raw_field": lambda: fx.FieldDetector(name="raw_field", grid_slice_tuple=full, exact_interpolation=False, components=("Ex", "Hz")),
        "field_reduce": lambda: fx.FieldDetector(name="field_reduce", grid_slice_tuple=plane, reduce_volume=True, components=("Ex", "Hy")),
        "energy": lambda: fx.EnergyDetector(name="energy", grid_slice_tuple=full, switch=fx.OnOffSwitch(interval=3)),
        "energy_slices": lambda: fx.EnergyDetector(name="energy_slices", grid_slice_tuple=full, as_slices=True, switch=fx.OnOffSwitch(interval=3)),
        "energy_pos": lambda: fx.EnergyDetector(name="energy_pos", grid_slice_tuple=full, as_slices=True, x_slice=2 * spacing, y_slice=3 * spacing, z_slice=4 * spacing),
        "energy_reduce": lambda: fx.EnergyDetector(name="energy_reduce", grid_slice_tuple=full, reduce_volume=True),
        "inverse_energy": lambda: fx.EnergyDetector(name="inverse_energy", grid_slice_tuple=full, as_slices=True, inverse=True, switch=fx.OnOffSwitch(interval=3)),
        "poynting": lambda: fx.PoyntingFluxDetector(name="poynting", grid_slice_tuple=plane, direction="+"),
        "poynting_interior": lambda: fx.PoyntingFluxDetector(name="poynting_interior", grid_slice_tuple=((off + 1, nx - off - 1), (off + 1, ny - off - 1), (nz - off - 3, nz - off - 2)), direction="+"),
        "energy_interior": lambda: fx.EnergyDetector(name="energy_interior", grid_slice_tuple=((off + 1, nx - off - 1), (off + 1, ny - off - 1), (off + 1, nz - off - 1)), reduce_volume=True),
        "poynting_full": lambda: fx.PoyntingFluxDetector(name="poynting_full", grid_slice_tuple=xplane, direction="-", reduce_volume=False),
        "poynting_all": lambda: fx.PoyntingFluxDetector(name="poynting_all", grid_slice_tuple=inner, direction="+", keep_all_components=True, fixed_propagation_axis=2),
        "phasor": lambda: fx.PhasorDetector(name="phasor", grid_slice_tuple=xplane, wave_characters=(wc, fx.WaveCharacter(wavelength=1.0e-6))),
        "phasor_pulse": lambda: fx.ModeOverlapDetector(name="phasor_pulse", grid_slice_tuple=plane, wave_characters=(wc,), scaling_mode="pulse", dft_subsample=2),
        "closed_flux": lambda: fx.ClosedSurfacePoyntingFluxDetector(name="closed_flux", grid_slice_tuple=((off, nx - off), (off, ny - off), (off, nz - off))),
        "closed_flux_in": lambda: fx.ClosedSurfacePoyntingFluxDetector(name="closed_flux_in", grid_slice_tuple=inner, orientation="inward", axes=(0, 2), switch=fx.OnOffSwitch(interval=2)),
        "closed_phasor": lambda: fx.ClosedSurfacePhasorPoyntingFluxDetector(name="closed_phasor", grid_slice_tuple=((off, nx - off), (off, ny - off), (off, nz - off)), wave_characters=(wc,)),
        "phasor_flux": lambda: fx.PhasorPoyntingFluxDetector(name="phasor_flux", grid_slice_tuple=plane, wave_characters=(wc, fx.WaveCharacter(wavelength=1.0e-6)), direction="+"),
        "phasor_reduce": lambda: fx.PhasorDetector(name="phasor_reduce", grid_slice_tuple=inner, wave_characters=(wc,), reduce_volume=True, components=("Ex", "Hy", "Hz")),
    }
    for d in detectors:
        objs.append(mk[d]())
    disp = None
    if poles:
        ct = coeff_tier
        disp = {
            "c1": (1.6 + 0.2 * rng.random((poles, ct, *shape))).astype(F),
            "c2": (-0.8 - 0.1 * rng.random((poles, ct, *shape))).astype(F),
            "c3": (0.05 * rng.random((poles, ct, *shape))).astype(F),
            "c4": (0.02 * rng.random((poles, ct, *shape))).astype(F) if c4 else None,
        }
    objects, arrays, _, cfg, _ = fx.place_objects(
        objs, cfg, inv_permittivities=inv_eps, inv_permeabilities=inv_mu, electric_conductivity=sE, magnetic_conductivity=sH, dispersive=disp
    )
    return objects, arrays, cfg


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a, b = a.astype(dt), b.astype(dt)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
