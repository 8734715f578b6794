"""Detector descriptors and state layout.

Mirrors ``fdtdx/objects/detectors/{detector,energy,field,poynting_flux,phasor,mode}.py`` as far
as the time step reads them: region, gate table ``on[t]``,
``arr_idx[t]``, interpolation flag, ``inverse`` flag, and the per-type reduction.  The per-step
``update`` arithmetic itself lives in the CUDA kernels (``csrc/detector_kernels.cuh``) and,
restated for checking, in ``oracle/yee.py``.  State arrays keep the reference's shapes/dtypes
(``detector.py:231-244``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Literal, Sequence

import numpy as np

from fdtdx_b200.boundaries import SimulationObject
from fdtdx_b200.switch import OnOffSwitch, WaveCharacter

_f32 = np.float32
COMPONENT_NAMES = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


@dataclass
class Detector(SimulationObject):
    dtype: type = np.float32
    exact_interpolation: bool = True
    inverse: bool = False
    switch: OnOffSwitch = field(default_factory=OnOffSwitch)
    _num_time_steps_on: int | None = None
    _is_on_at_time_step_arr: np.ndarray | None = None
    _time_step_to_arr_idx: np.ndarray | None = None
    _cached_cell_volume_weights: np.ndarray | None = None
    _dt: float = 0.0

    def _calculate_on_list(self, config) -> list[bool]:
        return self.switch.calculate_on_list(config.time_steps_total, config.time_step_duration)

    def _num_latent_time_steps(self) -> int:
        return int(self._num_time_steps_on)

    def place_on_grid(self, config):
        """``detector.py:195-229``."""
        self._dt = config.time_step_duration
        on_list = self._calculate_on_list(config)
        self._is_on_at_time_step_arr = np.asarray(on_list, dtype=bool)
        self._num_time_steps_on = int(sum(on_list))
        idx, counter = [-1] * len(on_list), 0
        for t, on in enumerate(on_list):
            if on:
                idx[t] = counter
                counter += 1
        self._time_step_to_arr_idx = np.asarray(idx, dtype=np.int32)
        grid = config.resolved_grid
        if grid is not None:
            self._cached_cell_volume_weights = grid.cell_volume(self.grid_slice_tuple)
        else:
            sp = config.uniform_spacing()
            self._cached_cell_volume_weights = (np.ones(self.grid_shape, _f32) * _f32(sp**3)).astype(_f32)
        return self

    def _shape_dtype_single_time_step(self) -> dict[str, tuple[tuple[int, ...], type]]:
        raise NotImplementedError

    def init_state(self) -> dict[str, np.ndarray]:
        n = self._num_latent_time_steps()
        return {k: np.zeros((n, *shape), dtype=dt) for k, (shape, dt) in self._shape_dtype_single_time_step().items()}


@dataclass
class EnergyDetector(Detector):
    """``energy.py``."""

    as_slices: bool = False
    reduce_volume: bool = False
    x_slice: float | None = None
    y_slice: float | None = None
    z_slice: float | None = None
    aggregate: str | None = None
    _slice_indices: tuple[int, int, int] | None = None

    @property
    def use_mean(self) -> bool:
        return self.aggregate == "mean" or any(s is None for s in (self.x_slice, self.y_slice, self.z_slice))

    def place_on_grid(self, config):
        super().place_on_grid(config)
        idxs = []
        for axis, pos in enumerate((self.x_slice, self.y_slice, self.z_slice)):
            n = self.grid_shape[axis]
            if pos is None:
                idxs.append(n // 2)
                continue
            grid = config.resolved_grid
            if grid is not None:
                lo, hi = self.grid_slice_tuple[axis]
                centers = grid.centers(axis)[lo:hi]
                idxs.append(int(np.clip(np.argmin(np.abs(centers - pos)), 0, n - 1)))
            else:
                sp = config.uniform_spacing()
                origin = self.grid_slice_tuple[axis][0] * sp
                idxs.append(max(0, min(int((pos - origin) / sp), n - 1)))
        self._slice_indices = tuple(idxs)
        return self

    def _shape_dtype_single_time_step(self):
        if self.as_slices and self.reduce_volume:
            raise Exception("Cannot both reduce volume and save slices!")
        gs = self.grid_shape
        if self.as_slices:
            return {
                "XY Plane": ((gs[0], gs[1]), self.dtype),
                "XZ Plane": ((gs[0], gs[2]), self.dtype),
                "YZ Plane": ((gs[1], gs[2]), self.dtype),
            }
        if self.reduce_volume:
            return {"energy": ((1,), self.dtype)}
        return {"energy": (gs, self.dtype)}


@dataclass
class FieldDetector(Detector):
    """``field.py``."""

    reduce_volume: bool = False
    components: Sequence[str] = COMPONENT_NAMES

    def _shape_dtype_single_time_step(self):
        n = len(self.components)
        return {"fields": ((n,) if self.reduce_volume else (n, *self.grid_shape), self.dtype)}


@dataclass
class PoyntingFluxDetector(Detector):
    """``poynting_flux.py:68-195``."""

    direction: Literal["+", "-"] = "+"
    reduce_volume: bool = True
    fixed_propagation_axis: int | None = None
    keep_all_components: bool = False
    _cached_face_area_weights: np.ndarray | None = None

    @property
    def propagation_axis(self) -> int:
        if self.fixed_propagation_axis is not None:
            if self.fixed_propagation_axis not in (0, 1, 2):
                raise Exception(f"Invalid: {self.fixed_propagation_axis=}")
            return self.fixed_propagation_axis
        if sum(a == 1 for a in self.grid_shape) != 1:
            raise Exception(f"Invalid poynting flux detector shape: {self.grid_shape}")
        return self.grid_shape.index(1)

    def place_on_grid(self, config):
        super().place_on_grid(config)
        grid = config.resolved_grid

        def weights(axis):
            if grid is not None:
                return grid.face_area(self.grid_slice_tuple, axis)
            sp = config.uniform_spacing()
            return (np.ones(self.grid_shape, _f32) * _f32(sp) * _f32(sp)).astype(_f32)

        if self.keep_all_components:
            self._cached_face_area_weights = np.stack([weights(a) for a in range(3)])
        else:
            self._cached_face_area_weights = weights(self.propagation_axis if grid is not None else 0)
        return self

    def _shape_dtype_single_time_step(self):
        if self.keep_all_components:
            shape = (3,) if self.reduce_volume else (3, *self.grid_shape)
        else:
            shape = (1,) if self.reduce_volume else self.grid_shape
        return {"poynting_flux": (shape, self.dtype)}


@dataclass
class PhasorDetector(Detector):
    """Running DFT (``phasor.py:21-235``).  State ``phasor``: ``(1, nf, nc, *region)`` complex64."""

    wave_characters: Sequence[WaveCharacter] = ()
    reduce_volume: bool = False
    components: Sequence[str] = COMPONENT_NAMES
    dtype: type = np.complex64
    scaling_mode: Literal["continuous", "pulse"] = "continuous"
    dft_subsample: int | str = 1
    _dft_stride: int = 1
    _window_at_time_step_arr: np.ndarray | None = None
    _window_sum: float | None = None

    def __post_init__(self):
        if self.dtype not in (np.complex64,):
            raise Exception(f"Invalid dtype in PhasorDetector: {self.dtype} (complex64 only on this backend)")

    @property
    def _angular_frequencies(self) -> np.ndarray:
        # 2 * jnp.pi * jnp.array(freqs): float32 array times the weak python scalar 2*pi
        return _f32(2 * np.pi) * np.array([wc.get_frequency() for wc in self.wave_characters], dtype=_f32)

    def _resolve_dft_stride(self, config) -> int:
        sub = self.dft_subsample
        if isinstance(sub, str):
            if sub != "auto":
                raise Exception(f"Invalid dft_subsample: {sub!r}")
            dt = float(config.time_step_duration)
            f_max = max((abs(float(wc.get_frequency())) for wc in self.wave_characters), default=0.0)
            if f_max <= 0.0 or dt <= 0.0:
                return 1
            return max(1, math.floor(1.0 / (8.0 * f_max * dt)))
        return max(1, int(sub))

    def _calculate_on_list(self, config) -> list[bool]:
        on_list = super()._calculate_on_list(config)
        stride = self._resolve_dft_stride(config)
        if stride <= 1:
            return on_list
        active = [t for t, on in enumerate(on_list) if on]
        kept = [False] * len(on_list)
        for t in active[::stride]:
            kept[t] = True
        return kept

    def place_on_grid(self, config):
        super().place_on_grid(config)
        self._dft_stride = self._resolve_dft_stride(config)
        window = self._is_on_at_time_step_arr.astype(_f32)
        self._window_at_time_step_arr = window
        self._window_sum = float(window.sum())
        if not math.isfinite(self._window_sum) or self._window_sum <= 0.0:
            raise Exception(f"Detector '{self.name}': the window sums to {self._window_sum}")
        return self

    def _static_scale(self) -> float | int:
        if self.scaling_mode == "continuous":
            return 2 / self._window_sum
        if self.scaling_mode == "pulse":
            return self._dft_stride
        raise Exception(f"Invalid scaling mode: {self.scaling_mode=}")

    def _num_latent_time_steps(self) -> int:
        return 1

    def _shape_dtype_single_time_step(self):
        nf, nc = len(self.wave_characters), len(self.components)
        gs = self.grid_shape if not self.reduce_volume else ()
        return {"phasor": ((nf, nc, *gs), np.complex64)}


def _face_area_weights(config, slice_tuple, axis: int) -> np.ndarray:
    """``_resolve_face_area_weights`` (``poynting_flux.py:14-43``)."""
    grid = config.resolved_grid
    if grid is not None:
        return grid.face_area(slice_tuple, axis)
    sp = config.uniform_spacing()
    shape = tuple(hi - lo for lo, hi in slice_tuple)
    return (np.ones(shape, _f32) * _f32(sp) * _f32(sp)).astype(_f32)


def _phasor_poynting_vector(phasors: np.ndarray) -> np.ndarray:
    """``Re(E x conj(H))`` of a (nf, 6, *spatial) phasor stack (``poynting_flux.py:58-70``)."""
    E, H = phasors[:, :3], np.conj(phasors[:, 3:])
    return np.stack([E[:, 1] * H[:, 2] - E[:, 2] * H[:, 1], E[:, 2] * H[:, 0] - E[:, 0] * H[:, 2], E[:, 0] * H[:, 1] - E[:, 1] * H[:, 0]], axis=1).real


@dataclass
class ClosedSurfacePoyntingFluxDetector(Detector):
    """Net Poynting flux through the faces of the detector box, one scalar per recorded step
    (``poynting_flux.py:199-284``, ``metrics.py:120-160``): for every active axis a,
    ``+sum(S_a * area)`` on the max face and ``-sum(S_a * area)`` on the min face."""

    orientation: Literal["outward", "inward"] = "outward"
    axes: tuple[int, ...] | None = None
    _face_area_weights_per_axis: np.ndarray | None = None

    def _resolve_active_axes(self) -> tuple[int, ...]:
        if self.axes is not None:
            return tuple(self.axes)
        return tuple(a for a in range(3) if self.grid_shape[a] > 1)

    def place_on_grid(self, config):
        if self.orientation not in ("outward", "inward"):
            raise ValueError(f"orientation must be 'outward' or 'inward', got {self.orientation!r}")
        if self.axes is not None and any(a not in (0, 1, 2) for a in self.axes):
            raise ValueError(f"axes entries must be in (0, 1, 2), got {self.axes}")
        super().place_on_grid(config)
        self._face_area_weights_per_axis = np.stack([_face_area_weights(config, self.grid_slice_tuple, a) for a in range(3)])
        return self

    def _shape_dtype_single_time_step(self):
        return {"poynting_flux": ((1,), self.dtype)}


def phasor_table(det: "PhasorDetector", T: int, dt: float) -> np.ndarray:
    """``exp(+i w t dt)`` for every time step and frequency as (T, nf, 2) float32 [cos, sin], rounded
    like ``phasor.py:186-214``: ``time_passed = t * dt`` and ``w * time_passed`` in float32."""
    tp = (np.arange(T, dtype=_f32) * _f32(dt)).astype(_f32)
    ang = (det._angular_frequencies[None, :] * tp[:, None]).astype(_f32)
    return np.ascontiguousarray(np.stack([np.cos(ang), np.sin(ang)], axis=-1), dtype=_f32)


@dataclass
class ModeOverlapDetector(PhasorDetector):
    """``objects/detectors/mode.py:193-470``: for the time step this *is* a six-component
    ``PhasorDetector``; ``apply`` solves the reference waveguide mode on the detector plane
    (``fdtdx_b200.modes``) and ``compute_overlap`` is the post-run overlap integral."""

    direction: Literal["+", "-"] = "+"
    mode_index: int = 0
    filter_pol: Literal["te", "tm"] | None = None
    _mode_E: np.ndarray | None = None
    _mode_H: np.ndarray | None = None
    _mode_neff: np.ndarray | None = None
    _cached_face_area_weights: np.ndarray | None = None

    def __post_init__(self):
        super().__post_init__()
        self.components = COMPONENT_NAMES

    @property
    def propagation_axis(self) -> int:
        if sum(a == 1 for a in self.grid_shape) != 1:
            raise Exception(f"Invalid ModeOverlapDetector shape: {self.grid_shape}")
        return self.grid_shape.index(1)

    def place_on_grid(self, config):
        super().place_on_grid(config)
        self._config = config
        if sum(a == 1 for a in self.grid_shape) == 1:
            w = _face_area_weights(config, self.grid_slice_tuple, self.propagation_axis)
            self._cached_face_area_weights = (w / w.mean()).astype(_f32)  # mode.py:246 (consistent with compute_mode)
        return self

    def apply(self, inv_permittivities, inv_permeabilities=1.0, inv_eps_slice=None):
        """Solve the reference mode per recorded frequency (``mode.py:302-377``).  ``inv_eps_slice``: the
        (C, *plane) cross-section itself, for callers that never materialise the volume on the host."""
        from fdtdx_b200 import modes

        gs = self.grid_slice
        inv_eps = np.asarray(inv_eps_slice) if inv_eps_slice is not None else np.asarray(inv_permittivities)[(slice(None), *gs)]
        mu = inv_permeabilities
        if hasattr(mu, "shape") and np.ndim(mu) > 0:
            mu = np.asarray(mu)[(slice(None), *gs)]
        cfg = self._config
        spacing = None if cfg.has_nonuniform_grid else cfg.uniform_spacing()
        Es, Hs, ns = [], [], []
        for wc in self.wave_characters:
            E, H, n = modes.compute_mode(wc.get_frequency(), inv_eps, mu, resolution=spacing, direction=self.direction, mode_index=self.mode_index,
                                         filter_pol=self.filter_pol, transverse_coords=modes._transverse_edges(cfg, self.grid_slice_tuple, self.propagation_axis))
            Es.append(E), Hs.append(H), ns.append(n)
        self._mode_E, self._mode_H, self._mode_neff = np.stack(Es), np.stack(Hs), np.asarray(ns)
        return self

    def compute_overlap(self, state) -> np.ndarray:
        """Complex overlap coefficient per frequency (``mode.py:379-470``)."""
        from fdtdx_b200 import modes

        ph = state["phasor"]
        ph = np.asarray(ph.cpu() if hasattr(ph, "cpu") else ph)[0]
        out = [modes.mode_overlap(ph[f], self._mode_E[f], self._mode_H[f], self.propagation_axis, self._cached_face_area_weights, pulse=self.scaling_mode == "pulse")
               for f in range(ph.shape[0])]
        return np.asarray(out)


@dataclass
class PhasorPoyntingFluxDetector(PhasorDetector):
    """Frequency-domain plane flux (``poynting_flux.py:287-388``): for the time step a six-component
    ``PhasorDetector``; ``compute_poynting_flux`` is the post-run surface integral."""

    direction: Literal["+", "-"] = "+"
    fixed_propagation_axis: int | None = None
    keep_all_components: bool = False
    _cached_face_area_weights: np.ndarray | None = None

    def __post_init__(self):
        super().__post_init__()
        self.components, self.reduce_volume = COMPONENT_NAMES, False

    @property
    def propagation_axis(self) -> int:
        if self.fixed_propagation_axis is not None:
            if self.fixed_propagation_axis not in (0, 1, 2):
                raise Exception(f"Invalid: {self.fixed_propagation_axis=}")
            return self.fixed_propagation_axis
        if sum(a == 1 for a in self.grid_shape) != 1:
            raise Exception(f"Invalid poynting flux detector shape: {self.grid_shape}")
        return self.grid_shape.index(1)

    def place_on_grid(self, config):
        super().place_on_grid(config)
        if self.keep_all_components:
            self._cached_face_area_weights = np.stack([_face_area_weights(config, self.grid_slice_tuple, a) for a in range(3)])
        else:
            self._cached_face_area_weights = _face_area_weights(config, self.grid_slice_tuple, self.propagation_axis)
        return self

    def compute_poynting_flux(self, state) -> np.ndarray:
        ph = np.asarray(state["phasor"].cpu() if hasattr(state["phasor"], "cpu") else state["phasor"])[0]
        pv = _phasor_poynting_vector(ph)
        if self.direction == "-":
            pv = -pv
        w = self._cached_face_area_weights
        flux = (pv * w[None]).sum(axis=(2, 3, 4)) if self.keep_all_components else (pv[:, self.propagation_axis] * w).sum(axis=(1, 2, 3))
        return 0.5 * flux if self.scaling_mode == "continuous" else flux


@dataclass
class ClosedSurfacePhasorPoyntingFluxDetector(PhasorDetector):
    """Frequency-domain closed-surface flux (``poynting_flux.py:391-540``): only the hollow shell is
    recorded - one six-component phasor plane per face of every active axis (state keys
    ``phasor_axis{a}_min`` / ``_max``); ``compute_net_flux`` integrates after the run.  The plan lowers
    it to one plane phasor accumulation per face."""

    orientation: Literal["outward", "inward"] = "outward"
    axes: tuple[int, ...] | None = None
    _face_area_weights_per_axis: tuple | None = None

    def __post_init__(self):
        super().__post_init__()
        self.components, self.reduce_volume = COMPONENT_NAMES, False

    def _resolve_active_axes(self) -> tuple[int, ...]:
        if self.axes is not None:
            return tuple(self.axes)
        return tuple(a for a in range(3) if self.grid_shape[a] > 1)

    def face_slices(self):
        """[(state key, grid_slice_tuple of the one-cell-thick face)] in state order."""
        out = []
        for a in self._resolve_active_axes():
            lo, hi = self.grid_slice_tuple[a]
            for side, (l_, h_) in (("min", (lo, lo + 1)), ("max", (hi - 1, hi))):
                sl = list(self.grid_slice_tuple)
                sl[a] = (l_, h_)
                out.append((f"phasor_axis{a}_{side}", tuple(sl)))
        return out

    def place_on_grid(self, config):
        if self.orientation not in ("outward", "inward"):
            raise ValueError(f"orientation must be 'outward' or 'inward', got {self.orientation!r}")
        if self.axes is not None and any(a not in (0, 1, 2) for a in self.axes):
            raise ValueError(f"axes entries must be in (0, 1, 2), got {self.axes}")
        super().place_on_grid(config)
        ws = []
        for a in range(3):
            w = _face_area_weights(config, self.grid_slice_tuple, a)
            idx = [slice(None)] * 3
            idx[a] = slice(0, 1)
            ws.append(w[tuple(idx)])
        self._face_area_weights_per_axis = tuple(ws)
        return self

    def _shape_dtype_single_time_step(self):
        nf = len(self.wave_characters)
        out = {}
        for key, sl in self.face_slices():
            out[key] = ((nf, 6, *[h_ - l_ for l_, h_ in sl]), np.complex64)
        return out

    def compute_net_flux(self, state) -> np.ndarray:
        net = np.zeros(len(self.wave_characters), np.float64)
        get = lambda k: np.asarray(state[k].cpu() if hasattr(state[k], "cpu") else state[k])[0]
        for a in self._resolve_active_axes():
            area = self._face_area_weights_per_axis[a]
            for side, sign in (("max", 1.0), ("min", -1.0)):
                s_a = _phasor_poynting_vector(get(f"phasor_axis{a}_{side}"))[:, a]
                net = net + sign * (s_a * area).sum(axis=(1, 2, 3))
        if self.orientation == "inward":
            net = -net
        return 0.5 * net if self.scaling_mode == "continuous" else net
