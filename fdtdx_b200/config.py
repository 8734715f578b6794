"""Host-side mirror of the reference's configuration types.

Mirrors ``fdtdx/config.py:15-307`` (``GradientConfig``, ``SimulationConfig``) and the parts of
``fdtdx/core/grid.py:284-560`` the Yee step reads (``UniformGrid``, ``RectilinearGrid``:
``cell_widths``, ``edges``, ``cfl_time_step``).  Only what the hot path consumes is restated;
the constraint solver and grid builders are out of scope.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Any, Literal, Sequence

import numpy as np

from fdtdx_b200 import constants


@dataclass(frozen=True)
class UniformGrid:
    """Uniform grid with one scalar spacing (reference ``core/grid.py`` ``UniformGrid``)."""

    spacing: float

    @property
    def is_uniform(self) -> bool:
        return True


class RectilinearGrid:
    """Rectilinear grid defined by its edge coordinates (``core/grid.py:284-560``).

    Edges are stored as float32 like the reference's default-dtype jax arrays, so that the
    metric scales derived from them (``curl.py:29-39``) round identically.
    """

    def __init__(self, x_edges: Sequence[float], y_edges: Sequence[float], z_edges: Sequence[float]):
        self.x_edges = np.asarray(x_edges, dtype=np.float32)
        self.y_edges = np.asarray(y_edges, dtype=np.float32)
        self.z_edges = np.asarray(z_edges, dtype=np.float32)
        self.dx = np.diff(self.x_edges)
        self.dy = np.diff(self.y_edges)
        self.dz = np.diff(self.z_edges)
        allw = np.concatenate([self.dx, self.dy, self.dz]).astype(np.float64)
        self._is_uniform = bool(np.allclose(allw, allw[0], rtol=1e-6, atol=0.0))
        self._uniform_spacing = float(allw[0]) if self._is_uniform else None

    @property
    def shape(self) -> tuple[int, int, int]:
        return (self.dx.shape[0], self.dy.shape[0], self.dz.shape[0])

    @property
    def is_uniform(self) -> bool:
        return self._is_uniform

    @property
    def uniform_spacing(self) -> float:
        if not self._is_uniform or self._uniform_spacing is None:
            raise ValueError("RectilinearGrid is non-uniform and has no single spacing")
        return self._uniform_spacing

    @property
    def min_spacings(self) -> tuple[float, float, float]:
        return (float(self.dx.min()), float(self.dy.min()), float(self.dz.min()))

    @property
    def min_spacing(self) -> float:
        return min(self.min_spacings)

    def edges(self, axis: int) -> np.ndarray:
        return (self.x_edges, self.y_edges, self.z_edges)[axis]

    def cell_widths(self, axis: int) -> np.ndarray:
        return (self.dx, self.dy, self.dz)[axis]

    def centers(self, axis: int) -> np.ndarray:
        e = self.edges(axis)
        return 0.5 * (e[:-1] + e[1:])

    def axis_extent(self, axis: int, bounds: tuple[int, int]) -> float:
        lower, upper = bounds
        e = self.edges(axis)
        return float(e[upper] - e[lower])

    def cell_volume(self, slice_tuple) -> np.ndarray:
        (x0, x1), (y0, y1), (z0, z1) = slice_tuple
        return (
            self.dx[x0:x1, None, None] * self.dy[None, y0:y1, None] * self.dz[None, None, z0:z1]
        ).astype(np.float32)

    def face_area(self, slice_tuple, axis: int) -> np.ndarray:
        """Face-area weights normal to ``axis`` on a grid slice (``poynting_flux.py`` helper)."""
        w = [self.dx, self.dy, self.dz]
        sl = [slice(*s) for s in slice_tuple]
        parts = []
        for a in range(3):
            shape = [1, 1, 1]
            n = slice_tuple[a][1] - slice_tuple[a][0]
            shape[a] = n
            parts.append(np.ones(n, np.float32).reshape(shape) if a == axis else w[a][sl[a]].reshape(shape))
        return (parts[0] * parts[1] * parts[2]).astype(np.float32)

    def cfl_time_step(self, courant_factor: float) -> float:
        """``core/grid.py:495-511``."""
        if self._is_uniform and self._uniform_spacing is not None:
            return (courant_factor / float(np.sqrt(3.0))) * self._uniform_spacing / constants.c
        dx_min, dy_min, dz_min = self.min_spacings
        inv_metric = (1 / dx_min**2) + (1 / dy_min**2) + (1 / dz_min**2)
        return courant_factor / (constants.c * float(np.sqrt(inv_metric)))


@dataclass(frozen=True)
class GradientConfig:
    """``config.py:15-53``."""

    method: Literal["reversible", "checkpointed"] = "reversible"
    recorder: Any = None
    num_checkpoints: int | None = None
    num_checkpoints_reversible: int = 0

    def __post_init__(self):
        if self.method == "reversible" and self.recorder is None:
            raise ValueError("reversible gradients need a Recorder")
        if self.method == "checkpointed" and self.num_checkpoints is None:
            raise ValueError("checkpointed gradients need num_checkpoints")


@dataclass(frozen=True)
class SimulationConfig:
    """``config.py:56-307``. ``backend`` is kept for signature parity; this backend is CUDA-only."""

    time: float
    grid: UniformGrid | RectilinearGrid
    backend: str = "gpu"
    dtype: Any = np.float32
    courant_factor: float = 0.99
    gradient_config: GradientConfig | None = None
    symmetry: tuple[int, int, int] = (0, 0, 0)
    use_complex_fields: bool | None = None  # None: complex iff a Bloch boundary has k != 0 (initialization.py:581-596)

    def aset(self, name: str, value: Any) -> "SimulationConfig":
        return replace(self, **{name: value})

    @property
    def courant_number(self) -> float:
        return self.courant_factor / math.sqrt(3)

    @property
    def resolved_grid(self) -> RectilinearGrid | None:
        return self.grid if isinstance(self.grid, RectilinearGrid) else None

    @property
    def has_nonuniform_grid(self) -> bool:
        g = self.resolved_grid
        return g is not None and not g.is_uniform

    def uniform_spacing(self) -> float:
        if isinstance(self.grid, UniformGrid):
            return self.grid.spacing
        return self.grid.uniform_spacing

    @property
    def time_step_duration(self) -> float:
        if isinstance(self.grid, RectilinearGrid):
            return self.grid.cfl_time_step(self.courant_factor)
        return self.courant_number * self.grid.spacing / constants.c

    @property
    def time_steps_total(self) -> int:
        return round(self.time / self.time_step_duration)

    @property
    def max_travel_distance(self) -> float:
        return constants.c * self.time

    @property
    def only_forward(self) -> bool:
        return self.gradient_config is None

    @property
    def invertible_optimization(self) -> bool:
        return self.gradient_config is not None and self.gradient_config.recorder is not None
