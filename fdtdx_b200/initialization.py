"""Array allocation for a placed scene (stands in for ``fdtdx/fdtd/initialization.py``).

The reference's ``place_objects`` (constraint solver + rasterisation, ``initialization.py:100-315``)
is OUT OF SCOPE.  What the hot path needs from it is the set of arrays
``_init_arrays`` allocates (``initialization.py:552-842, 1141-1165``) with the reference's shapes
and dtypes; this module allocates exactly those for objects that already carry explicit grid
slices, plus a small box rasteriser for isotropic / diagonal / full-tensor materials.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np

from fdtdx_b200.boundaries import PerfectlyMatchedLayer, SimulationObject, SimulationVolume
from fdtdx_b200.config import SimulationConfig
from fdtdx_b200.container import ArrayContainer, FieldState, ObjectContainer, RecordingState
from fdtdx_b200.detectors import Detector
from fdtdx_b200.sources import Source

_f32 = np.float32


@dataclass
class Material:
    """Subset of ``fdtdx/materials.py`` ``Material``: scalar, 3-tuple (diagonal) or 9-tuple
    (row-major full tensor) permittivity / permeability / conductivities."""

    permittivity: Any = 1.0
    permeability: Any = 1.0
    electric_conductivity: Any = 0.0
    magnetic_conductivity: Any = 0.0


@dataclass
class UniformMaterialObject(SimulationObject):
    material: Material = field(default_factory=Material)


def _tier(v) -> int:
    if np.isscalar(v):
        return 1
    return len(v)


def _as_tensor9(v) -> np.ndarray:
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    if v.size == 1:
        return np.diag([v[0]] * 3).reshape(9)
    if v.size == 3:
        return np.diag(v).reshape(9)
    return v.reshape(9)


def _inv_components(v, tier: int) -> np.ndarray:
    """Inverse material components for a global tier (1, 3 or 9), ``initialization.py:713-740``."""
    if tier == 9:
        return np.linalg.inv(_as_tensor9(v).reshape(3, 3)).reshape(9)
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    if tier == 3:
        v = np.repeat(v, 3) if v.size == 1 else v
    return 1.0 / v


def _components(v, tier: int) -> np.ndarray:
    if tier == 9:
        return _as_tensor9(v)
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    if tier == 3 and v.size == 1:
        v = np.repeat(v, 3)
    return v


def rasterize_materials(volume_shape, config: SimulationConfig, background: Material, objects: Sequence[UniformMaterialObject]):
    """Paint box materials in order; returns (inv_eps, inv_mu, sigma_E, sigma_H) in the reference's
    conventions: inv_mu is the python float 1.0 when nothing is magnetic; sigma arrays are None
    when nothing conducts and are stored pre-multiplied by ``c0*dt/courant``."""
    from fdtdx_b200.constants import c as c0

    mats = [background] + [o.material for o in objects]
    eps_tier = max(max(_tier(m.permittivity), _tier(m.electric_conductivity)) for m in mats)
    mu_tier = max(max(_tier(m.permeability), _tier(m.magnetic_conductivity)) for m in mats)
    magnetic = any(not (np.isscalar(m.permeability) and m.permeability == 1.0) for m in mats) or mu_tier > 1
    cond_E = any(np.any(np.asarray(m.electric_conductivity) != 0) for m in mats)
    cond_H = any(np.any(np.asarray(m.magnetic_conductivity) != 0) for m in mats)
    ref_spacing = c0 * config.time_step_duration / config.courant_number

    def paint(get, tier, inverse):
        arr = np.empty((tier, *volume_shape), _f32)
        fn = _inv_components if inverse else _components
        arr[:] = fn(get(background), tier).astype(_f32)[:, None, None, None]
        for o in objects:
            arr[(slice(None), *o.grid_slice)] = fn(get(o.material), tier).astype(_f32)[:, None, None, None]
        return arr

    inv_eps = paint(lambda m: m.permittivity, eps_tier, True)
    inv_mu = paint(lambda m: m.permeability, mu_tier, True) if magnetic else 1.0
    sigma_E = paint(lambda m: m.electric_conductivity, eps_tier, False) * _f32(ref_spacing) if cond_E else None
    sigma_H = paint(lambda m: m.magnetic_conductivity, mu_tier, False) * _f32(ref_spacing) if cond_H else None
    return inv_eps, inv_mu, sigma_E, sigma_H


def init_arrays(
    objects: ObjectContainer,
    config: SimulationConfig,
    inv_permittivities: np.ndarray,
    inv_permeabilities: Any = 1.0,
    electric_conductivity: np.ndarray | None = None,
    magnetic_conductivity: np.ndarray | None = None,
    dispersive: dict | None = None,
) -> ArrayContainer:
    """Allocate the step's arrays with the reference's layouts (``initialization.py:552-842``)."""
    shape = objects.volume.grid_shape
    # complex-valued fields only when a Bloch boundary carries a non-zero wave vector (initialization.py:581-596)
    needs_complex = any(getattr(b, "needs_complex_fields", False) for b in objects.boundary_objects)
    use_complex = needs_complex if config.use_complex_fields is None else bool(config.use_complex_fields)
    if needs_complex and not use_complex:
        raise ValueError(
            "use_complex_fields=False but Bloch boundaries with non-zero wave vector are present. These require complex-valued fields."
        )
    fdt = np.complex64 if use_complex else _f32
    E = np.zeros((3, *shape), fdt)
    H = np.zeros((3, *shape), fdt)
    psi_E, psi_H = {}, {}
    for pml in objects.pml_objects:
        psi_E[pml.name] = (np.zeros(pml.grid_shape, fdt), np.zeros(pml.grid_shape, fdt))
        psi_H[pml.name] = (np.zeros(pml.grid_shape, fdt), np.zeros(pml.grid_shape, fdt))
    det_states = {d.name: d.init_state() for d in objects.detectors}
    rec_state = None
    gc = config.gradient_config
    if gc is not None and gc.recorder is not None:
        rec = gc.recorder.init_tables(config.time_steps_total)
        import torch

        data = {}
        for pml in objects.pml_objects:
            for fs in ("E", "H"):
                shp = (rec._latent_array_size, 3, *pml.interface_grid_shape())
                if rec.dtype_code == 0:
                    data[f"{pml.name}_{fs}"] = np.zeros(shp, fdt)
                elif use_complex:
                    raise NotImplementedError("DtypeConversion of complex (Bloch) interface recordings")
                else:
                    from fdtdx_b200.container import _TorchLeaf

                    data[f"{pml.name}_{fs}"] = _TorchLeaf(torch.zeros(shp, dtype=rec.torch_dtype()))
        rec_state = RecordingState(data=data, state={})
    P_curr = P_prev = c1 = c2 = c3 = c4 = None
    if dispersive is not None:
        c1, c2, c3 = (np.asarray(dispersive[k], _f32) for k in ("c1", "c2", "c3"))
        c4 = None if dispersive.get("c4") is None else np.asarray(dispersive["c4"], _f32)
        npoles = c1.shape[0]
        P_curr = np.zeros((npoles, 3, *shape), _f32)
        P_prev = np.zeros((npoles, 3, *shape), _f32)
    return ArrayContainer(
        fields=FieldState(E=E, H=H, psi_E=psi_E, psi_H=psi_H, dispersive_P_curr=P_curr, dispersive_P_prev=P_prev),
        inv_permittivities=np.asarray(inv_permittivities, _f32),
        inv_permeabilities=inv_permeabilities if np.isscalar(inv_permeabilities) else np.asarray(inv_permeabilities, _f32),
        detector_states=det_states,
        recording_state=rec_state,
        electric_conductivity=None if electric_conductivity is None else np.asarray(electric_conductivity, _f32),
        magnetic_conductivity=None if magnetic_conductivity is None else np.asarray(magnetic_conductivity, _f32),
        dispersive_c1=c1,
        dispersive_c2=c2,
        dispersive_c3=c3,
        dispersive_c4=c4,
    )


def place_objects(
    object_list: Sequence[SimulationObject],
    config: SimulationConfig,
    constraints: Sequence | None = None,
    key: Any = None,
    *,
    inv_permittivities: np.ndarray | None = None,
    inv_permeabilities: Any = None,
    electric_conductivity: np.ndarray | None = None,
    magnetic_conductivity: np.ndarray | None = None,
    dispersive: dict | None = None,
    background: Material | None = None,
):
    """Signature-compatible entry (``initialization.py:100-111``) for *already placed* objects.

    ``constraints`` must be empty/None (the constraint DSL is out of scope).  Materials come either
    from explicit arrays or from ``UniformMaterialObject`` boxes over ``background``.
    Returns ``(objects, arrays, params, config, info)`` like the reference; ``params`` is ``{}``.
    """
    if constraints:
        raise NotImplementedError("the placement-constraint solver is outside the hot-path scope; give explicit grid slices")
    object_list = list(object_list)
    vol_idx = next(i for i, o in enumerate(object_list) if isinstance(o, SimulationVolume))
    shape = object_list[vol_idx].grid_shape
    for o in object_list:
        if isinstance(o, (Source, Detector)) and getattr(o, "_is_on_at_time_step_arr", None) is None:
            o.place_on_grid(config)
        if isinstance(o, PerfectlyMatchedLayer) and o.pml_a_E is None:
            o.place_on_grid(config)
    objects = ObjectContainer(object_list=object_list, volume_idx=vol_idx)
    if inv_permittivities is None:
        boxes = [o for o in object_list if isinstance(o, UniformMaterialObject)]
        inv_eps, inv_mu, sE, sH = rasterize_materials(shape, config, background or Material(), boxes)
        inv_permittivities = inv_eps
        inv_permeabilities = inv_mu if inv_permeabilities is None else inv_permeabilities
        electric_conductivity = sE if electric_conductivity is None else electric_conductivity
        magnetic_conductivity = sH if magnetic_conductivity is None else magnetic_conductivity
    if inv_permeabilities is None:
        inv_permeabilities = 1.0
    arrays = init_arrays(objects, config, inv_permittivities, inv_permeabilities, electric_conductivity, magnetic_conductivity, dispersive)
    return objects, arrays, {}, config, {}
