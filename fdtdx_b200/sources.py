"""Source objects: the per-step injection state the Yee step reads.

Mirrors ``fdtdx/objects/sources/{source,tfsf,linear_polarization,mode,dipole}.py`` as far as the
hot path goes: a TFSF plane source is, for the time step, the four
arrays ``_E, _H, _time_offset_E, _time_offset_H`` of shape ``(3, *face)`` plus a temporal
profile, a direction and a switch.  The profile *builders* below (uniform plane, tilted Gaussian
beam, synthetic mode) are simplified host-side setup helpers - the reference's ``apply()``
(profile construction, tidy3d mode solve) is out of scope and marked "next" (section 8 f1).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Literal

import numpy as np

from fdtdx_b200 import constants
from fdtdx_b200.boundaries import SimulationObject
from fdtdx_b200.profile import SingleFrequencyProfile
from fdtdx_b200.switch import OnOffSwitch, WaveCharacter

_f32 = np.float32


def get_oriented_transverse_axes(normal_axis: int) -> tuple[int, int]:
    """``core/axis.py:20-23``."""
    return ((normal_axis + 1) % 3, (normal_axis + 2) % 3)


@dataclass
class Source(SimulationObject):
    wave_character: WaveCharacter = None  # type: ignore[assignment]
    temporal_profile: object = field(default_factory=SingleFrequencyProfile)
    static_amplitude_factor: float = 1.0
    switch: OnOffSwitch = field(default_factory=OnOffSwitch)
    _is_on_at_time_step_arr: np.ndarray | None = None
    _time_step_to_on_idx: np.ndarray | None = None

    def _update_on_arrays(self, config):
        """``source.py:51-66``."""
        T, dt = config.time_steps_total, config.time_step_duration
        self._is_on_at_time_step_arr = np.asarray(self.switch.calculate_on_list(T, dt), dtype=bool)
        self._time_step_to_on_idx = np.asarray(self.switch.calculate_time_step_to_on_arr_idx(T, dt), dtype=np.int32)
        return self

    def place_on_grid(self, config):
        return self._update_on_arrays(config)

    @property
    def uses_default_switch(self) -> bool:
        return self.switch.is_default_always_on

    def adjusted_time_step(self, time_step: int) -> np.float32:
        """``source.py:44-49`` -> ``linear_interpolated_indexing`` at an integer point
        (``core/misc.py:363-391``): both corner weights are 1, so the result is
        ``2*v / (2 + 1e-8)`` which is exactly ``v`` in float32."""
        return _f32(self._time_step_to_on_idx[time_step])


@dataclass
class TFSFPlaneSource(Source):
    """``tfsf.py:412-806``: one-cell-thick plane, direction +/-, per-cell Yee time offsets."""

    direction: Literal["+", "-"] = "+"
    _E: np.ndarray | None = None
    _H: np.ndarray | None = None
    _time_offset_E: np.ndarray | None = None
    _time_offset_H: np.ndarray | None = None
    _temporal_H_filter: np.ndarray | None = None

    @property
    def propagation_axis(self) -> int:
        gs = self.grid_shape
        if sum(a == 1 for a in gs) < 1:
            raise Exception(f"Invalid plane source shape: {gs}")
        return gs.index(1)

    def metric_scale_at_plane(self, config, stencil: str) -> float:
        """``tfsf.py:699-737``."""
        if not config.has_nonuniform_grid:
            return 1.0
        grid = config.resolved_grid
        n = self.propagation_axis
        start = self.grid_slice_tuple[n][0]
        widths = grid.cell_widths(n)
        width = widths[start]
        if stencil == "backward":
            width = _f32(0.5) * (width + widths[max(start - 1, 0)])
        ref = constants.c * config.time_step_duration / config.courant_number
        return float(_f32(ref) / _f32(width))


def _yee_positions(n: int, offset: float, spacing=None, edges=None) -> np.ndarray:
    if edges is None:
        edges = np.arange(n + 1, dtype=_f32) * _f32(spacing)
    if offset == 0:
        return edges[:-1]
    return _f32(0.5) * (edges[:-1] + edges[1:])


def calculate_time_offset_yee(center_physical, wave_vector, refractive_idx, face_shape, config, slice_tuple):
    """Per-component Yee time offsets of a plane wave on a face (``core/grid.py:720-880``).

    offset[q, cell] = -(r_q(cell) - center) . k_hat * n / (c * dt), with r_q the Yee position of
    component q.  Host-side setup; returns two float32 arrays of shape (3, *face_shape).
    """
    grid = config.resolved_grid
    coords = {}
    for ax in range(3):
        n = face_shape[ax]
        if grid is not None:
            lo, hi = slice_tuple[ax]
            e = grid.edges(ax)[lo : hi + 1].astype(_f32) - grid.edges(ax)[lo].astype(_f32)
            coords[ax] = (_yee_positions(n, 0, edges=e), _yee_positions(n, 0.5, edges=e))
        else:
            sp = config.uniform_spacing()
            coords[ax] = (_yee_positions(n, 0, spacing=sp), _yee_positions(n, 0.5, spacing=sp))

    def xyz(offsets):
        c = [coords[ax][1 if offsets[ax] == 0.5 else 0] for ax in range(3)]
        x, y, z = np.meshgrid(c[0], c[1], c[2], indexing="ij")
        return np.stack([x, y, z], axis=-1) - np.asarray(center_physical, dtype=_f32)[None, None, None, :]

    wv = np.asarray(wave_vector, dtype=_f32)
    xyz_E = np.stack([xyz((0.5, 0, 0)), xyz((0, 0.5, 0)), xyz((0, 0, 0.5))])
    xyz_H = np.stack([xyz((0, 0.5, 0.5)), xyz((0.5, 0, 0.5)), xyz((0.5, 0.5, 0))])
    velocity = (_f32(constants.c) / np.asarray(refractive_idx, dtype=_f32))[None, ...]
    tE = (-(xyz_E @ wv)) / (velocity * _f32(config.time_step_duration))
    tH = (-(xyz_H @ wv)) / (velocity * _f32(config.time_step_duration))
    return tE.astype(_f32), tH.astype(_f32)


def _polarization_vectors(direction, axis, e_pol, azimuth=0.0, elevation=0.0):
    """Wave vector and (E, H) polarisation of a (possibly tilted) plane wave.

    Untilted: k = +/- e_axis, H = k x E.  Tilt rotates k about the vertical (azimuth) and
    horizontal (elevation) transverse axes and re-orthogonalises E (cf. ``tfsf.py`` /
    ``tilted_polarization_vectors``).
    """
    h_ax, v_ax = get_oriented_transverse_axes(axis)
    k = np.zeros(3)
    s = 1.0 if direction == "+" else -1.0
    k[axis] = s * np.cos(azimuth) * np.cos(elevation)
    k[h_ax] = np.sin(azimuth) * np.cos(elevation)
    k[v_ax] = np.sin(elevation)
    k /= np.linalg.norm(k)
    e = np.asarray(e_pol, dtype=np.float64)
    e = e - np.dot(e, k) * k
    e /= np.linalg.norm(e)
    h = np.cross(k, e)
    return e.astype(_f32), h.astype(_f32), k.astype(_f32)


def make_plane_source(
    name: str,
    grid_slice_tuple,
    config,
    inv_permittivities: np.ndarray,
    inv_permeabilities=1.0,
    *,
    direction: Literal["+", "-"] = "+",
    wave_character: WaveCharacter,
    temporal_profile=None,
    fixed_E_polarization_vector=(1.0, 0.0, 0.0),
    amplitude_profile: np.ndarray | None = None,
    azimuth_angle: float = 0.0,
    elevation_angle: float = 0.0,
    normalize_by_energy: bool = True,
    static_amplitude_factor: float = 1.0,
    switch: OnOffSwitch | None = None,
    effective_index: float | None = None,
    dispersive: dict | None = None,
) -> TFSFPlaneSource:
    """Uniform / Gaussian / mode-like plane source builder (setup helper, float32).

    ``amplitude_profile`` is a transverse weight of the face's shape (default: ones = uniform
    plane wave).  E = w * e_pol, H = w * h_pol / impedance; optional energy normalisation; per-cell
    time offsets from the local refractive index (or ``effective_index`` for mode sources).
    """
    src = TFSFPlaneSource(
        name=name,
        grid_slice_tuple=grid_slice_tuple,
        wave_character=wave_character,
        temporal_profile=temporal_profile or SingleFrequencyProfile(),
        static_amplitude_factor=static_amplitude_factor,
        switch=switch or OnOffSwitch(),
        direction=direction,
    )
    src.place_on_grid(config)
    axis = src.propagation_axis
    face = src.grid_shape
    gs = src.grid_slice
    inv_eps = np.asarray(inv_permittivities)[:, gs[0], gs[1], gs[2]].astype(_f32)
    inv_eps_inf = inv_eps
    disp = None
    if dispersive is not None and dispersive.get("c1") is not None:
        # linear_polarization.py:195-213: impedance / normalisation / time offsets use the real effective
        # permittivity at the carrier frequency, not eps_infinity
        from fdtdx_b200 import dispersion as _disp

        disp = {k: (None if dispersive.get(k) is None else np.asarray(dispersive[k])[(slice(None), slice(None), *gs)]) for k in ("c1", "c2", "c3", "c4")}
        inv_eps = _disp.effective_inv_permittivity(
            inv_eps, disp["c1"], disp["c2"], disp["c3"], 2.0 * np.pi * wave_character.get_frequency(), config.time_step_duration, disp["c4"]
        ).astype(_f32)
    if inv_eps.shape[0] == 9:
        inv_eps_iso = inv_eps[0]
    else:
        inv_eps_iso = inv_eps[0]
    if isinstance(inv_permeabilities, np.ndarray) and inv_permeabilities.ndim > 0:
        inv_mu_iso = inv_permeabilities[:, gs[0], gs[1], gs[2]][0].astype(_f32)
    else:
        inv_mu_iso = np.full(face, _f32(inv_permeabilities), _f32)
    e_pol, h_pol, k = _polarization_vectors(
        direction, axis, fixed_E_polarization_vector, np.deg2rad(azimuth_angle), np.deg2rad(elevation_angle)
    )
    w = np.ones(face, _f32) if amplitude_profile is None else np.asarray(amplitude_profile, _f32).reshape(face)
    E = w[None] * e_pol[:, None, None, None]
    H = w[None] * h_pol[:, None, None, None]
    if normalize_by_energy:
        energy = _f32(0.5) * ((E * E).sum(0) / inv_eps_iso + (H * H).sum(0) / inv_mu_iso)
        root = np.sqrt(energy.sum(dtype=_f32))
        E, H = E / root, H / root
    impedance = np.sqrt(inv_eps_iso / inv_mu_iso)
    H = H / impedance[None]
    # centre of the face in physical coordinates (metres from the slice's lower corner)
    grid = config.resolved_grid
    center = []
    for ax in range(3):
        if ax == axis:
            center.append(0.0)
        elif grid is not None:
            lo, hi = grid_slice_tuple[ax]
            e = grid.edges(ax)
            center.append(0.5 * float(e[hi] - e[lo]))
        else:
            center.append(0.5 * (face[ax] - 1) * config.uniform_spacing())
    n_idx = (
        np.full(face, _f32(effective_index), _f32)
        if effective_index is not None
        else (_f32(1.0) / np.sqrt(inv_eps_iso * inv_mu_iso)).astype(_f32)
    )
    tE, tH = calculate_time_offset_yee(center, k, n_idx, face, config, grid_slice_tuple)
    src._E, src._H = E.astype(_f32), H.astype(_f32)
    src._time_offset_E, src._time_offset_H = tE, tH
    if disp is not None:
        # broadband impedance correction (linear_polarization.py:330-351, tfsf.py:21-140): the H-side
        # amplitude becomes a table sampled at integer time steps, interpolated at t + offset
        from fdtdx_b200 import dispersion as _disp

        T = config.time_steps_total
        times = (np.arange(T, dtype=np.float64) * config.time_step_duration).astype(_f32)
        raw = src.temporal_profile.get_amplitude(time=times, period=wave_character.get_period(), phase_shift=wave_character.phase_shift)
        filt = _disp.dispersive_H_filter(
            raw, config.time_step_duration, disp["c1"], disp["c2"], disp["c3"], inv_eps_inf,
            2.0 * np.pi * wave_character.get_frequency(), disp["c4"],
        )
        src._temporal_H_filter = np.asarray(filt, _f32)
    return src


def gaussian_amplitude_profile(face_shape, axis: int, radius_cells: float, std: float = 1 / 3) -> np.ndarray:
    """Radially symmetric Gaussian weight exp(-r^2 / (2 (std*R)^2)) cut at r > R (cf.
    ``GaussianPlaneSource._get_amplitude_raw``)."""
    h_ax, v_ax = get_oriented_transverse_axes(axis)
    idx = np.indices(face_shape).astype(np.float64)
    ch, cv = 0.5 * (face_shape[h_ax] - 1), 0.5 * (face_shape[v_ax] - 1)
    r2 = (idx[h_ax] - ch) ** 2 + (idx[v_ax] - cv) ** 2
    w = np.exp(-0.5 * r2 / (std * radius_cells) ** 2)
    w[r2 > radius_cells**2] = 0.0
    return w.astype(_f32)


@dataclass
class PointDipoleSource(Source):
    """``dipole.py:195-277`` (axis-aligned electric or magnetic point dipole)."""

    polarization: int = 2
    amplitude: float = 1.0
    source_type: Literal["electric", "magnetic"] = "electric"
