"""Synthetic builders for the BASELINE.json configurations (used by bench.py, smoke and tests).

* ``build_coupler`` - C2, ``performance/directional_coupler.py:39-260`` of the reference: SOI
  directional coupler on the script's non-uniform grid (``make_grid`` :84-109), 12-cell CPML,
  SiO2 cladding + 220 nm Si cores, x-normal TE0 mode plane with a Gaussian pulse, three
  mode-overlap (phasor) planes with pulse scaling.  The reference rasterises ``coupler.gds`` with
  gdstk/matplotlib and solves the mode with tidy3d; neither is available here, so the core
  geometry is generated analytically (two 0.5 um arms, 236 nm gap over 20 um, cosine S-bends to
  the four ports, port stubs to the domain edge) and the mode profile is a synthetic TE0-like
  Gaussian fed identically to oracle and kernels ("mode-profile parity
  unpinned").  ``n_copies`` chains that many couplers along x (weak scaling: one per rank).
* ``build_box`` - C5: vacuum with a centred eps=2.25 box of half the side, 10-cell CPML, a dipole.

Arrays are created directly on ``device`` (torch) so that billion-cell scenes never touch the host;
``device=None`` gives NumPy arrays for the oracle.
"""

from __future__ import annotations

import math

import numpy as np

import fdtdx_b200 as fx
from fdtdx_b200.container import ArrayContainer, FieldState

_f32 = np.float32
EPS_SI, EPS_SIO2 = 12.25, 2.25


def coupler_grid(cells_per_lambda: int, n_copies: int = 1, domain=(42e-6, 7e-6, 4e-6), wavelength=1550e-9):
    """``make_grid`` (directional_coupler.py:84-109), x repeated ``n_copies`` times."""
    n_si, n_sio2 = math.sqrt(EPS_SI), math.sqrt(EPS_SIO2)
    dx_f = wavelength / (n_si * cells_per_lambda)
    dx_c = wavelength / (n_sio2 * cells_per_lambda)
    dom_x, dom_y, dom_z = domain
    nx1 = round(dom_x / dx_f)
    x_edges = np.linspace(0.0, dom_x * n_copies, nx1 * n_copies + 1)
    y_edges = np.concatenate(
        [
            np.linspace(0.0, 0.5e-6, round(0.5e-6 / dx_c) + 1),
            np.linspace(0.5e-6, dom_y - 0.5e-6, round((dom_y - 1e-6) / dx_f) + 1)[1:],
            np.linspace(dom_y - 0.5e-6, dom_y, round(0.5e-6 / dx_c) + 1)[1:],
        ]
    )
    z_edges = np.concatenate(
        [
            np.linspace(0.0, 1.0e-6, round(1.0e-6 / dx_c) + 1),
            np.linspace(1.0e-6, 3.0e-6, round(2.0e-6 / dx_f) + 1)[1:],
            np.linspace(3.0e-6, dom_z, round(1.0e-6 / dx_c) + 1)[1:],
        ]
    )
    return fx.RectilinearGrid(x_edges, y_edges, z_edges), nx1, dx_f


def coupler_core_mask(grid, nx1: int, x0: int, x1: int):
    """Boolean (x1-x0, Ny) core mask + (Nz,) z mask of the SOI coupler for the global x range [x0, x1).

    Geometry: the layer-1 polygons of the reference's ``performance/coupler.gds`` (extracted once by
    ``scripts/extract_coupler_gds.py`` into ``fdtdx_b200/data/coupler_gds.npz``) plus the four port stubs
    that ``extend_gds_with_port_stubs`` adds (``directional_coupler.py:117-140``): 0.5 um wide
    rectangles from each port to the domain edge; Si where a cell CENTRE lies inside; 220 nm thick
    around the mid-plane (``CouplerConfig.si_z_base``)."""
    import os

    from fdtdx_b200.gds import rasterize_polygons

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "coupler_gds.npz"))
    polys = [d[f"p{i}"] for i in range(int(d["n"]))]
    dom_x = float(grid.x_edges[nx1])
    gds_center = (10e-6, 0.368e-6)
    left, right = gds_center[0] - dom_x / 2, gds_center[0] + dom_x / 2
    for px, py, orient in ((-10e-6, -1.632e-6, 180.0), (-10e-6, 2.368e-6, 180.0), (30e-6, 2.368e-6, 0.0), (30e-6, -1.632e-6, 0.0)):
        xa, xb = (left, px) if orient == 180.0 else (px, right)
        polys.append(np.array([[xa, py - 0.25e-6], [xb, py - 0.25e-6], [xb, py + 0.25e-6], [xa, py + 0.25e-6]]))
    xc = grid.centers(0)[x0:x1].astype(np.float64)
    yc = grid.centers(1).astype(np.float64)
    zc = grid.centers(2).astype(np.float64)
    gx = np.mod(xc, dom_x) - dom_x / 2 + gds_center[0]  # sim -> GDS frame (chained copies repeat)
    gy = yc - 3.5e-6 + gds_center[1]
    in_xy = rasterize_polygons(polys, gx, gy) if np.all(np.diff(gx) > 0) else np.concatenate(
        [rasterize_polygons(polys, gx[a:b], gy) for a, b in _monotone_runs(gx)], axis=0)
    in_z = (zc >= 2e-6 - 110e-9) & (zc <= 2e-6 + 110e-9)
    return in_xy, in_z


def _monotone_runs(v):
    cuts = [0] + [i + 1 for i in range(len(v) - 1) if v[i + 1] <= v[i]] + [len(v)]
    return list(zip(cuts[:-1], cuts[1:]))


def build_coupler(cells_per_lambda: int = 20, device=None, n_copies: int = 1, x_range=None, with_detectors=True, time_steps_cap=None):
    """C2 (``performance/directional_coupler.py:183-260``): returns (objects, arrays, config).  With
    ``x_range=(x0, x1)`` only that x-slab of every array is allocated (objects keep global
    coordinates; the plan clips them).  The TE0 mode of the o1 port cross-section (2 um x full z) drives
    an x-normal mode plane with a Gaussian pulse; three ``ModeOverlapDetector`` planes (pulse scaling) sit
    one fine cell after the source and at the thru / cross ports, each with its solved reference mode."""
    from fdtdx_b200.modes import make_mode_source

    grid, nx1, dx_f = coupler_grid(cells_per_lambda, n_copies)
    shape = grid.shape
    sim_time = 2.0 * math.sqrt(EPS_SI) * 42e-6 / 3e8
    cfg = fx.SimulationConfig(time=sim_time, grid=grid, gradient_config=None)
    nx, ny, nz = shape
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=((0, nx), (0, ny), (0, nz)))
    bl = fx.boundary_objects_from_config(shape, cfg, "pml", thickness=12)
    x0, x1 = x_range if x_range is not None else (0, nx)
    in_xy, in_z = coupler_core_mask(grid, nx1, x0, x1)
    wc = fx.WaveCharacter(wavelength=1550e-9)
    prof = fx.GaussianPulseProfile(center_wave=wc, spectral_width=fx.WaveCharacter(wavelength=1550e-9 * 10))
    objs = [vol, *bl]
    xe = grid.x_edges.astype(np.float64)

    def y_window(center_y):
        lo = int(np.searchsorted(grid.y_edges, center_y - 1e-6))
        hi = int(np.searchsorted(grid.y_edges, center_y + 1e-6))
        return lo, max(hi, lo + 1)

    def plane_inv_eps(ix, ylo, yhi):
        """(1, 1, ny_window, Nz) inverse permittivity of the cross-section at global plane ix: every
        rank solves the same mode, so the plane is rasterised from the geometry, not read from a slab."""
        m_xy, m_z = coupler_core_mask(grid, nx1, ix, ix + 1)
        core = m_xy[0, ylo:yhi, None] & m_z[None, :]
        return np.where(core, _f32(1.0 / EPS_SI), _f32(1.0 / EPS_SIO2)).astype(_f32)[None, None]

    ports = {"o1": (-10e-6, -1.632e-6), "o2": (-10e-6, 2.368e-6), "o3": (30e-6, 2.368e-6), "o4": (30e-6, -1.632e-6)}
    to_sim = lambda p: (p[0] + 21e-6 - 10e-6, p[1] + 3.5e-6 - 0.368e-6)
    mode_cache = {}
    for c in range(n_copies):
        off = c * nx1
        sx, sy = to_sim(ports["o1"])
        ix = off + int(np.searchsorted(xe[: nx1 + 1], sx))
        ylo, yhi = y_window(sy)
        sl = ((ix, ix + 1), (ylo, yhi), (0, nz))
        key = ("src", ix - off, ylo, yhi)
        if key not in mode_cache:
            mode_cache[key] = make_mode_source("tmp", sl, cfg, None, direction="+", wave_character=wc, temporal_profile=prof, mode_index=0, filter_pol="te",
                                               inv_eps_slice=plane_inv_eps(ix, ylo, yhi))
        m = mode_cache[key]
        src = fx.sources.TFSFPlaneSource(name=f"source{c}" if n_copies > 1 else "source", grid_slice_tuple=sl, wave_character=wc, temporal_profile=prof, direction="+")
        src.place_on_grid(cfg)
        src._E, src._H, src._time_offset_E, src._time_offset_H, src._neff = m._E, m._H, m._time_offset_E, m._time_offset_H, m._neff
        objs.append(src)
        if with_detectors:
            for name, port, shift in (("det_source", "o1", 1), ("det_thru", "o4", 0), ("det_cross", "o3", 0)):
                px_, py_ = to_sim(ports[port])
                dix = off + int(np.searchsorted(xe[: nx1 + 1], px_)) + shift
                dlo, dhi = y_window(py_)
                dsl = ((dix, dix + 1), (dlo, dhi), (0, nz))
                det = fx.ModeOverlapDetector(name=f"{name}{c}" if n_copies > 1 else name, grid_slice_tuple=dsl, wave_characters=(wc,), scaling_mode="pulse",
                                             direction="+", mode_index=0, filter_pol="te")
                det.place_on_grid(cfg)
                key = ("det", dix - off, dlo, dhi)
                if key not in mode_cache:
                    det.apply(None, inv_eps_slice=plane_inv_eps(dix, dlo, dhi))
                    mode_cache[key] = (det._mode_E, det._mode_H, det._mode_neff)
                det._mode_E, det._mode_H, det._mode_neff = mode_cache[key]
                objs.append(det)
    objects = fx.ObjectContainer(objs)
    arrays = _alloc(objects, cfg, (x0, x1), device, lambda: _coupler_eps(in_xy, in_z, device))
    return objects, arrays, cfg


def _coupler_eps(in_xy, in_z, device):
    if device is None:
        core = in_xy[:, :, None] & in_z[None, None, :]
        return np.where(core, _f32(1.0 / EPS_SI), _f32(1.0 / EPS_SIO2)).astype(_f32)[None]
    import torch

    a = torch.from_numpy(in_xy).to(device)
    b = torch.from_numpy(in_z).to(device)
    core = a[:, :, None] & b[None, None, :]
    out = torch.where(core, torch.tensor(1.0 / EPS_SI, dtype=torch.float32, device=device), torch.tensor(1.0 / EPS_SIO2, dtype=torch.float32, device=device))
    return out[None].contiguous()


def _alloc(objects, cfg, x_range, device, eps_fn):
    """Allocate the step arrays for the x-slab ``x_range`` on ``device`` (None -> NumPy)."""
    shape = objects.volume.grid_shape
    x0, x1 = x_range
    local = (x1 - x0, shape[1], shape[2])
    if device is None:
        z = lambda *s, dtype=_f32: np.zeros(s, dtype)
    else:
        import torch

        tdt = {_f32: torch.float32, np.complex64: torch.complex64}
        z = lambda *s, dtype=_f32: torch.zeros(*s, dtype=tdt.get(dtype, torch.float32), device=device)
    psi_E, psi_H = {}, {}
    for pml in objects.pml_objects:
        gs = list(pml.grid_slice_tuple)
        lo, hi = max(gs[0][0], x0), min(gs[0][1], x1)
        if hi <= lo:
            continue
        shp = (hi - lo, gs[1][1] - gs[1][0], gs[2][1] - gs[2][0])
        psi_E[pml.name] = (z(*shp), z(*shp))
        psi_H[pml.name] = (z(*shp), z(*shp))
    det_states = {}
    for d in objects.detectors:
        lo, hi = d.grid_slice_tuple[0]
        if lo >= x0 and hi <= x1:
            st = d.init_state()
            det_states[d.name] = {k: (v if device is None else z(*v.shape, dtype=v.dtype.type)) for k, v in st.items()}
    return ArrayContainer(
        fields=FieldState(E=z(3, *local), H=z(3, *local), psi_E=psi_E, psi_H=psi_H),
        inv_permittivities=eps_fn(),
        inv_permeabilities=1.0,
        detector_states=det_states,
        recording_state=None,
    )


def build_box(shape=(512, 512, 512), device=None, x_range=None, spacing=50e-9, thickness=10, time=1e-12):
    """C5: vacuum + centred eps=2.25 box of half the side, CPML on all faces, one dipole."""
    cfg = fx.SimulationConfig(time=time, grid=fx.UniformGrid(spacing=spacing))
    nx, ny, nz = shape
    vol = fx.SimulationVolume(name="volume", grid_slice_tuple=((0, nx), (0, ny), (0, nz)))
    bl = fx.boundary_objects_from_config(shape, cfg, "pml", thickness=thickness)
    wc = fx.WaveCharacter(wavelength=20 * spacing)
    c = (nx // 2, ny // 2, nz // 5)
    dip = fx.PointDipoleSource(name="dipole", grid_slice_tuple=tuple((v, v + 1) for v in c), wave_character=wc, polarization=0, amplitude=1.0)
    dip.place_on_grid(cfg)
    objects = fx.ObjectContainer([vol, *bl, dip])
    x0, x1 = x_range if x_range is not None else (0, nx)

    def eps_fn():
        bx = (nx // 4, nx - nx // 4)
        by = (ny // 4, ny - ny // 4)
        bz = (nz // 4, nz - nz // 4)
        lo, hi = max(bx[0], x0) - x0, max(min(bx[1], x1) - x0, 0)
        if device is None:
            e = np.ones((1, x1 - x0, ny, nz), _f32)
        else:
            import torch

            e = torch.ones((1, x1 - x0, ny, nz), dtype=torch.float32, device=device)
        if hi > lo:
            e[0, lo:hi, by[0] : by[1], bz[0] : bz[1]] = 1.0 / 2.25
        return e

    arrays = _alloc(objects, cfg, (x0, x1), device, eps_fn)
    return objects, arrays, cfg


def bytes_per_cell_step(objects, arrays, local_shape) -> float:
    """Algorithmic HBM bytes per cell per full step: each field read once and
    written once per half-step, material arrays read once, CPML psi read+written in slab cells."""
    n_eps = int(arrays.inv_permittivities.shape[0])
    mu = arrays.inv_permeabilities
    n_mu = int(mu.shape[0]) if hasattr(mu, "shape") and len(mu.shape) > 0 else 0
    b = 2 * (12 + 12 + 12) + 4 * n_eps + 4 * n_mu
    if arrays.electric_conductivity is not None:
        b += 4 * int(arrays.electric_conductivity.shape[0])
    if arrays.magnetic_conductivity is not None:
        b += 4 * int(arrays.magnetic_conductivity.shape[0])
    if arrays.dispersive_c1 is not None:
        npol = int(arrays.dispersive_c1.shape[0])
        ct = int(arrays.dispersive_c1.shape[1])
        b += npol * (36 + 4 * ct * (3 if arrays.dispersive_c4 is None else 4))
    cells = float(np.prod(local_shape))
    psi_cells = 0.0
    for name, (a, _) in arrays.fields.psi_E.items():
        psi_cells += float(np.prod(a.shape))
    b += 32.0 * psi_cells / cells
    return float(b)
