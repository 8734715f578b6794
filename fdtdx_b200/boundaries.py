"""Boundary objects the Yee step consumes.

Mirrors ``fdtdx/objects/boundaries/{boundary,perfectly_matched_layer,pec,pmc,bloch}.py``.
Only the hooks the time step reads are restated: ``axis``, ``direction``,
``grid_slice_tuple``, ``uses_wrap_padding``, the CPML coefficient tables and the interface slice.
Objects carry explicit grid slices; the reference's constraint solver is out of scope.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Literal

import numpy as np

from fdtdx_b200.constants import c, eps0, eta0

SliceTuple3D = tuple[tuple[int, int], tuple[int, int], tuple[int, int]]


@dataclass
class SimulationObject:
    name: str = ""
    grid_slice_tuple: SliceTuple3D = ((0, 0), (0, 0), (0, 0))

    @property
    def grid_slice(self) -> tuple[slice, slice, slice]:
        return tuple(slice(a, b) for a, b in self.grid_slice_tuple)  # type: ignore[return-value]

    @property
    def grid_shape(self) -> tuple[int, int, int]:
        return tuple(b - a for a, b in self.grid_slice_tuple)  # type: ignore[return-value]


@dataclass
class SimulationVolume(SimulationObject):
    """The whole domain; ``grid_shape`` is the only thing the step reads (``update.py:128``)."""


@dataclass
class BaseBoundary(SimulationObject):
    axis: int = 0
    direction: Literal["+", "-"] = "-"
    _is_symmetry_wall: bool = False

    @property
    def uses_wrap_padding(self) -> bool:
        return False

    @property
    def thickness(self) -> int:
        return 1

    def interface_slice_tuple(self) -> SliceTuple3D:
        """``boundary.py:117-131``: first slab cell for '+', last slab cell for '-'."""
        s = [*self.grid_slice_tuple]
        lo, hi = self.grid_slice_tuple[self.axis]
        s[self.axis] = (lo, lo + 1) if self.direction == "+" else (hi - 1, hi)
        return (s[0], s[1], s[2])

    def interface_slice(self):
        return tuple(slice(a, b) for a, b in self.interface_slice_tuple())

    def interface_grid_shape(self):
        g = list(self.grid_shape)
        g[self.axis] = 1
        return tuple(g)


@dataclass
class PerfectlyMatchedLayer(BaseBoundary):
    """CPML slab (``perfectly_matched_layer.py:12-303``)."""

    alpha_start: float | None = None
    alpha_end: float | None = None
    alpha_order: float | None = None
    kappa_start: float | None = None
    kappa_end: float | None = None
    kappa_order: float | None = None
    sigma_start: float | None = None
    sigma_end: float | None = None
    sigma_order: float | None = None
    pml_a_E: np.ndarray | None = None
    pml_b_E: np.ndarray | None = None
    inv_kappa_E: np.ndarray | None = None
    pml_a_H: np.ndarray | None = None
    pml_b_H: np.ndarray | None = None
    inv_kappa_H: np.ndarray | None = None

    def __post_init__(self):
        # defaults: perfectly_matched_layer.py:69-95
        if self.alpha_start is None:
            self.alpha_start = 0.01 * 2 * np.pi * c / 1.55e-6 * eps0
        if self.alpha_end is None:
            self.alpha_end = 0.0
        if self.alpha_order is None:
            self.alpha_order = 1.0
        if self.kappa_start is None:
            self.kappa_start = 1.0
        if self.kappa_end is None:
            self.kappa_end = 1.0
        if self.kappa_order is None:
            self.kappa_order = 3.0
        if self.sigma_start is None:
            self.sigma_start = 0.0
        if self.sigma_order is None:
            self.sigma_order = 3.0

    @property
    def thickness(self) -> int:
        return self.grid_shape[self.axis]

    def _physical_thickness(self, config) -> float:
        grid = config.resolved_grid
        if grid is not None:
            return grid.axis_extent(self.axis, self.grid_slice_tuple[self.axis])
        return self.thickness * config.uniform_spacing()

    def place_on_grid(self, config) -> "PerfectlyMatchedLayer":
        """Build the a/b/1/kappa tables (``perfectly_matched_layer.py:97-136``), float32."""
        f32 = np.float32
        if self.sigma_end is None:
            L_phys = self._physical_thickness(config)
            # the reference computes this with float32 jnp.log and casts to float
            self.sigma_end = float(f32(-(self.sigma_order + 1)) * np.log(f32(1e-6)) / f32(2 * (eta0 / 1.0) * L_phys))
        dt = config.time_step_duration
        sigma_E, sigma_H = self._compute_pml_profile(config, self.sigma_start, self.sigma_end, self.sigma_order)
        kappa_E, kappa_H = self._compute_pml_profile(config, self.kappa_start, self.kappa_end, self.kappa_order)
        alpha_E, alpha_H = self._compute_pml_profile(config, self.alpha_start, self.alpha_end, self.alpha_order)

        def coeffs(sigma, kappa, alpha):
            with np.errstate(invalid="ignore", divide="ignore"):
                b = (np.expm1(f32(-dt / eps0) * (sigma / kappa + alpha)) + f32(1)).astype(f32)
                a = ((b - f32(1.0)) * sigma / (sigma + alpha * kappa) / kappa).astype(f32)
            a = np.where(np.isnan(a), f32(0.0), a).astype(f32)
            return a, b

        self.pml_a_E, self.pml_b_E = coeffs(sigma_E, kappa_E, alpha_E)
        self.pml_a_H, self.pml_b_H = coeffs(sigma_H, kappa_H, alpha_H)
        self.inv_kappa_E = (f32(1.0) / kappa_E).astype(f32)
        self.inv_kappa_H = (f32(1.0) / kappa_H).astype(f32)
        return self

    def _compute_pml_profile(self, config, value_start, value_end, order):
        """``perfectly_matched_layer.py:231-275`` (+ non-uniform depths ``:277-303``)."""
        f32 = np.float32
        L = self.thickness
        if config.has_nonuniform_grid:
            grid = config.resolved_grid
            lower, upper = self.grid_slice_tuple[self.axis]
            edges = grid.edges(self.axis)[lower : upper + 1].astype(f32)
            norm = float(edges[-1] - edges[0])
            centers = f32(0.5) * (edges[:-1] + edges[1:])
            zero = np.zeros(1, f32)
            if self.direction == "-":
                interface = edges[-1]
                dE = interface - edges[1:]
                dH = np.concatenate([interface - centers[1:], zero])
            else:
                interface = edges[0]
                dE = np.concatenate([zero, centers[:-1] - interface])
                dH = edges[:-1] - interface
        elif self.direction == "-":
            dE = np.arange(L - 1, -1, -1, dtype=f32)
            dH = np.append(np.arange(L - 1.5, -0.5, -1, dtype=f32), f32(0))
            norm = L
        else:
            dE = np.insert(np.arange(0.5, L - 0.5, 1, dtype=f32), 0, f32(0))
            dH = np.arange(0, L, 1, dtype=f32)
            norm = L
        pE = (f32(value_start) + f32(value_end - value_start) * np.power(dE / f32(norm), f32(order))).astype(f32)
        pH = (f32(value_start) + f32(value_end - value_start) * np.power(dH / f32(norm), f32(order))).astype(f32)
        shape = [1, 1, 1]
        shape[self.axis] = L
        return pE.reshape(shape), pH.reshape(shape)

    @property
    def kappa_is_one(self) -> bool:
        return self.kappa_start == 1.0 and self.kappa_end == 1.0


@dataclass
class PerfectElectricConductor(BaseBoundary):
    """``pec.py``: zero tangential E on the 1-cell wall after every E update."""

    @property
    def tangential_components(self) -> tuple[int, int]:
        return {0: (1, 2), 1: (0, 2), 2: (0, 1)}[self.axis]


@dataclass
class PerfectMagneticConductor(BaseBoundary):
    """``pmc.py``: zero tangential H on the 1-cell wall after every H update."""

    @property
    def tangential_components(self) -> tuple[int, int]:
        return {0: (1, 2), 1: (0, 2), 2: (0, 1)}[self.axis]


@dataclass
class BlochBoundary(BaseBoundary):
    """``objects/boundaries/bloch.py``: ``F(x + L) = F(x) exp(i k L)``.  A zero Bloch vector is the plain
    periodic boundary; a non-zero component along this boundary's axis needs
    complex fields (``initialization.py:581-596``), which the kernels run as two real systems whose
    wrapped ghost planes are mixed with the phase (``fdtdx_b200/bloch.py``)."""

    bloch_vector: tuple[float, float, float] = (0.0, 0.0, 0.0)

    @property
    def uses_wrap_padding(self) -> bool:
        return True

    @property
    def needs_complex_fields(self) -> bool:
        """``bloch.py:31-38``: only the component along this boundary's axis matters."""
        return self.bloch_vector[self.axis] != 0.0

    def get_bloch_phase(self, volume_shape, config):
        """``bloch.py:135-155``: ``exp(i k_axis L)`` as complex64, ``L`` = physical extent of the axis.
        jnp evaluates the weakly typed ``1j * k * L`` in float32, so the angle is rounded first."""
        k = self.bloch_vector[self.axis]
        if config.has_nonuniform_grid:
            edges = np.asarray(config.resolved_grid.edges(self.axis), dtype=np.float32)
            L = np.float32(edges[volume_shape[self.axis]] - edges[0])
        else:
            L = volume_shape[self.axis] * config.uniform_spacing()
        theta = np.float32(k * L)
        return np.complex64(complex(np.cos(theta), np.sin(theta)))


PeriodicBoundary = BlochBoundary


def boundary_objects_from_config(
    volume_shape: tuple[int, int, int],
    config,
    types: dict[str, str] | str = "pml",
    thickness: int = 10,
    bloch_vector: tuple[float, float, float] = (0.0, 0.0, 0.0),
) -> list[BaseBoundary]:
    """Build the six boundary objects of a box (stands in for the reference's
    ``BoundaryConfig`` + ``boundary_objects_from_config``, ``objects/boundaries/initialization.py``).

    ``types`` maps "min_x".."max_z" to "pml" | "periodic" | "bloch" | "pec" | "pmc" (or one string for all);
    "bloch" faces carry ``bloch_vector`` (rad/m).
    """
    names = ["min_x", "max_x", "min_y", "max_y", "min_z", "max_z"]
    if isinstance(types, str):
        types = {n: types for n in names}
    out: list[BaseBoundary] = []
    for n in names:
        kind = types[n]
        axis = "xyz".index(n[-1])
        direction = "-" if n.startswith("min") else "+"
        t = thickness if kind == "pml" else 1
        sl = [(0, volume_shape[0]), (0, volume_shape[1]), (0, volume_shape[2])]
        sl[axis] = (0, t) if direction == "-" else (volume_shape[axis] - t, volume_shape[axis])
        kw = dict(name=f"{kind}_{n}", grid_slice_tuple=tuple(sl), axis=axis, direction=direction)
        if kind == "pml":
            out.append(PerfectlyMatchedLayer(**kw).place_on_grid(config))
        elif kind == "periodic":
            out.append(BlochBoundary(**kw))
        elif kind == "bloch":
            out.append(BlochBoundary(bloch_vector=tuple(float(v) for v in bloch_vector), **kw))
        elif kind == "pec":
            out.append(PerfectElectricConductor(**kw))
        elif kind == "pmc":
            out.append(PerfectMagneticConductor(**kw))
        else:
            raise ValueError(f"unknown boundary type {kind!r}")
    return out
