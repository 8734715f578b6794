"""BASELINE.json configurations at their real sizes and step counts.

Each builder returns ``(objects, arrays, config)`` with NumPy arrays (oracle side); the GPU tests move
them with ``arrays.to_torch("cuda")``.  Geometry, grid, boundary thickness, source kind, detector set,
recorder modules and simulated time follow the reference example scripts line by line; what the
reference derives with out-of-scope machinery (constraint solver, tidy3d mode solve, parameter
transforms) is replaced by explicit grid slices, the in-repo mode solver / a seeded design block - noted
per builder.

C1  examples/simulate_gaussian_source.py:35-179        120^3, 525 steps, bf16 recorder, 2 energy videos
C3a examples/dispersive_gaussian_pulse.py:45-180       (3,3,900), 9179 steps, Lorentz pole, filtered-H source
C3b examples/simulate_gaussian_source_fully_anisotropic.py:35-133   120^3, full-tensor slab, f32 recorder
C4  examples/optimize_ceviche_corner.py:45-251         (135,135,75), 1311 steps, every-5 + fp8 recorder
C2s performance/directional_coupler.py at a reduced cells-per-lambda (the bench scene itself)
"""

from __future__ import annotations

import numpy as np

import fdtdx_b200 as fx
from fdtdx_b200 import dispersion as disp

F = np.float32


def _box(shape):
    return ((0, shape[0]), (0, shape[1]), (0, shape[2]))


def build_c1(pml: bool = True, time: float = 100e-15):
    """Tilted Gaussian beam in eps = 2, two full-volume energy videos (one inverse), bf16 recorder.
    ``pml=True`` is the script's ``periodic = False`` branch (10-cell CPML, the variant with recorded
    interfaces); ``pml=False`` the shipped periodic one."""
    shape = (120, 120, 120)
    rec = fx.Recorder(modules=[fx.DtypeConversion(dtype="bfloat16")])
    cfg = fx.SimulationConfig(time=time, grid=fx.UniformGrid(spacing=100e-9), courant_factor=0.99,
                              gradient_config=fx.GradientConfig(recorder=rec))
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=_box(shape))
    bl = fx.boundary_objects_from_config(shape, cfg, "pml" if pml else "periodic", thickness=10)
    inv_eps = np.full((1, *shape), 1.0 / 2.0, F)
    sl = ((10, 110), (10, 110), (60, 61))  # 10 um x 10 um face centred in the volume
    amp = fx.gaussian_amplitude_profile((100, 100, 1), 2, radius_cells=40.0, std=1 / 3)
    src = fx.make_plane_source("source", sl, cfg, inv_eps, 1.0, direction="-", wave_character=fx.WaveCharacter(wavelength=1.55e-6),
                               fixed_E_polarization_vector=(1.0, 0.0, 0.0), amplitude_profile=amp, elevation_angle=-20.0)
    dets = [
        fx.EnergyDetector(name="Energy Video", grid_slice_tuple=_box(shape), as_slices=True, switch=fx.OnOffSwitch(interval=3)),
        fx.EnergyDetector(name="Backwards Energy Video", grid_slice_tuple=_box(shape), as_slices=True, switch=fx.OnOffSwitch(interval=3), inverse=True),
    ]
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl, src, *dets], cfg, inv_permittivities=inv_eps)
    return objects, arrays, cfg


# dispersive_gaussian_pulse.py:47-66
C3A = dict(eps_inf=7.98737492, omega_0=3.93282466e15, delta_eps=3.68799143, gamma=1e8, sigma_t=4.0e-15, res=20e-9, pml=10,
           source_z=12, det_a=100, det_b=800, time=350e-15, wavelength=1.55e-6)


def build_c3a(time: float | None = None, nz: int = 900):
    """Gaussian pulse through Lorentz-dispersive silicon: ADE pole everywhere, periodic x/y, CPML z,
    uniform plane source with the broadband H-side filter table, two plane-mean Ex traces."""
    p = C3A
    shape = (3, 3, nz)
    cfg = fx.SimulationConfig(time=time or p["time"], grid=fx.UniformGrid(spacing=p["res"]), courant_factor=0.99)
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=_box(shape))
    types = {"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}
    bl = fx.boundary_objects_from_config(shape, cfg, types, thickness=p["pml"])
    inv_eps = np.full((1, *shape), 1.0 / p["eps_inf"], F)
    pole = disp.LorentzPole(resonance_frequency=p["omega_0"], damping=p["gamma"], delta_epsilon=p["delta_eps"])
    coeffs = disp.coefficient_arrays((pole,), cfg.time_step_duration, np.ones(shape, bool))
    wc = fx.WaveCharacter(wavelength=p["wavelength"])
    prof = fx.GaussianPulseProfile(spectral_width=fx.WaveCharacter(frequency=1.0 / (2.0 * np.pi * p["sigma_t"])), center_wave=wc)
    sl = ((0, 3), (0, 3), (p["source_z"], p["source_z"] + 1))
    src = fx.make_plane_source("source", sl, cfg, inv_eps, 1.0, direction="+", wave_character=wc, temporal_profile=prof,
                               fixed_E_polarization_vector=(1.0, 0.0, 0.0), dispersive=coeffs)
    dets = []
    for name, z in (("pulse_trace_A", p["det_a"]), ("pulse_trace_B", min(p["det_b"], nz - p["pml"] - 2))):
        dets.append(fx.FieldDetector(name=name, grid_slice_tuple=((0, 3), (0, 3), (z, z + 1)), components=("Ex",), reduce_volume=True))
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl, src, *dets], cfg, inv_permittivities=inv_eps, dispersive=coeffs)
    return objects, arrays, cfg


def build_c3b(time: float = 100e-15, n: int = 120):
    """Plane wave onto a fully anisotropic slab: nine-component inv_eps, periodic x/y, CPML z,
    full-volume Ey video every 3rd step, float32 recorder (no modules)."""
    shape = (n, n, 120)
    rec = fx.Recorder(modules=[])
    cfg = fx.SimulationConfig(time=time, grid=fx.UniformGrid(spacing=100e-9), courant_factor=0.99, gradient_config=fx.GradientConfig(recorder=rec))
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=_box(shape))
    types = {"min_x": "periodic", "max_x": "periodic", "min_y": "periodic", "max_y": "periodic", "min_z": "pml", "max_z": "pml"}
    bl = fx.boundary_objects_from_config(shape, cfg, types, thickness=10)
    slab = fx.UniformMaterialObject(name="slab", grid_slice_tuple=((0, n), (0, n), (50, 70)),
                                    material=fx.Material(permittivity=(2.5, 1.5, 0.0, 1.5, 2.5, 0.0, 0.0, 0.0, 1.0)))
    from fdtdx_b200.initialization import rasterize_materials

    inv_eps, inv_mu, _, _ = rasterize_materials(shape, cfg, fx.Material(), [slab])
    sl = ((0, n), (0, n), (100, 101))  # centre + 4 um
    src = fx.make_plane_source("source", sl, cfg, inv_eps[(0, 4, 8), ...], 1.0, direction="-", wave_character=fx.WaveCharacter(wavelength=1.55e-6),
                               fixed_E_polarization_vector=(1.0, 0.0, 0.0))
    det = fx.FieldDetector(name="Electric Field Video", grid_slice_tuple=_box(shape), components=("Ey",), switch=fx.OnOffSwitch(interval=3))
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl, slab, src, det], cfg, inv_permittivities=inv_eps, inv_permeabilities=inv_mu)
    return objects, arrays, cfg


def c4_geometry():
    """Grid slices of optimize_ceviche_corner.py resolved by hand (dx = 20 nm, 10-cell CPML):
    substrate z in [0,25); device 80x80x20 cells flush with the max-x / min-y CPML faces minus the
    0.2 um margins; input guide along x into the device, output guide along y out of it."""
    shape = (135, 135, 75)
    dev = ((135 - 10 - 10 - 80, 135 - 10 - 10), (10 + 10, 10 + 10 + 80), (25, 45))
    wg_in = ((0, dev[0][0]), ((dev[1][0] + dev[1][1]) // 2 - 10, (dev[1][0] + dev[1][1]) // 2 + 10), (25, 45))
    wg_out = (((dev[0][0] + dev[0][1]) // 2 - 10, (dev[0][0] + dev[0][1]) // 2 + 10), (dev[1][1], 135), (25, 45))
    return shape, dev, wg_in, wg_out


def c4_inv_eps(seed: int = 0):
    shape, dev, wg_in, wg_out = c4_geometry()
    eps = np.ones(shape, np.float64)
    eps[:, :, 0:25] = 2.25  # silica substrate
    for b in (wg_in, wg_out):
        eps[b[0][0]:b[0][1], b[1][0]:b[1][1], b[2][0]:b[2][1]] = 12.25
    # seeded continuous design: uniform[0,1] parameters on the 80x80 voxel grid, smoothed by a fixed 2-D
    # Gaussian (std 3 voxels, the script's GaussianSmoothing2D) and mapped linearly onto eps in [1, 12.25]
    rng = np.random.default_rng(seed)
    par = rng.random((80, 80))
    k = np.exp(-0.5 * (np.arange(-9, 10) / 3.0) ** 2)
    k /= k.sum()
    sm = np.apply_along_axis(lambda v: np.convolve(np.pad(v, 9, mode="edge"), k, mode="valid"), 0, par)
    sm = np.apply_along_axis(lambda v: np.convolve(np.pad(v, 9, mode="edge"), k, mode="valid"), 1, sm)
    eps[dev[0][0]:dev[0][1], dev[1][0]:dev[1][1], dev[2][0]:dev[2][1]] = (1.0 + 11.25 * sm)[:, :, None]
    return (1.0 / eps).astype(F)[None]


def build_c4(time: float = 50e-15, seed: int = 0, gradient: bool = True):
    """Corner-bend inverse design step: CW TE0 mode source (in-repo mode solver), two Poynting planes
    gated to one optical period each, last-step energy slices, recorder [every-5, fp8_e4m3fnuz]."""
    from fdtdx_b200.modes import make_mode_source

    shape, dev, wg_in, wg_out = c4_geometry()
    gc = fx.GradientConfig(recorder=fx.Recorder(modules=[fx.LinearReconstructEveryK(k=5), fx.DtypeConversion(dtype="float8_e4m3fnuz")])) if gradient else None
    cfg = fx.SimulationConfig(time=time, grid=fx.UniformGrid(spacing=20e-9), courant_factor=0.99, gradient_config=gc)
    T = cfg.time_steps_total
    wl = 1.55e-6
    period_steps = round((wl / fx.constants.c) / cfg.time_step_duration)
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=_box(shape))
    bl = fx.boundary_objects_from_config(shape, cfg, "pml", thickness=10)
    inv_eps = c4_inv_eps(seed)
    sx = 10 + 4  # grid_margins = thickness + 4 from the guide's low-x end
    src = make_mode_source("source", ((sx, sx + 1), (0, 135), (0, 75)), cfg, inv_eps, direction="+", wave_character=fx.WaveCharacter(wavelength=wl), mode_index=0, filter_pol="te")
    steps = list(range(T))
    on_in = steps[7 * period_steps:8 * period_steps] if T >= 8 * period_steps else steps[-period_steps:]
    flux_in = fx.PoyntingFluxDetector(name="in flux", grid_slice_tuple=((sx + 3, sx + 4), (0, 135), (0, 75)), direction="+", switch=fx.OnOffSwitch(fixed_on_time_steps=on_in))
    oy = 135 - 10 - 5 - 1
    flux_out = fx.PoyntingFluxDetector(name="out flux", grid_slice_tuple=((0, 135), (oy, oy + 1), (0, 75)), direction="+", switch=fx.OnOffSwitch(fixed_on_time_steps=steps[-period_steps:]))
    e_last = fx.EnergyDetector(name="energy_last_step", grid_slice_tuple=_box(shape), as_slices=True, switch=fx.OnOffSwitch(fixed_on_time_steps=[-1]))
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl, src, flux_in, flux_out, e_last], cfg, inv_permittivities=inv_eps)
    return objects, arrays, cfg


def build_c2(cells_per_lambda: int = 6, max_steps: int | None = None):
    """The bench scene (fdtdx_b200.workloads.build_coupler) at a resolution the CPU oracle can run."""
    from fdtdx_b200 import workloads as W

    objects, arrays, cfg = W.build_coupler(cells_per_lambda, device=None)
    return objects, arrays, cfg


# ----------------------------------------------------------------------------------------------
# what the config-parity goldens keep of a run (shared by tests/golden/make_config_golden.py, which
# fills them from the CPU oracle, and tests/test_config_parity.py, which fills them from the CUDA path)
# ----------------------------------------------------------------------------------------------
def sub(a, stride=4):
    """Every ``stride``-th cell of the last three axes (full-size fields are tens of MB)."""
    a = np.asarray(a)
    return np.ascontiguousarray(a[..., ::stride, ::stride, ::stride])


def field_norms(a):
    """Per-component L2 norms over the WHOLE array in float64: a size-independent checksum of the
    cells the subsample skips."""
    a = np.asarray(a, np.float64)
    return np.sqrt((a.reshape(a.shape[0], -1) ** 2).sum(axis=1))


def t_pick(n, k=6):
    """k time indices spread over [0, n) including the last."""
    return np.unique(np.linspace(0, n - 1, k).round().astype(int))
