"""Array/object containers: the boundary types of the drop-in seam.

Mirrors ``fdtdx/fdtd/container.py`` (``ObjectContainer:33``, ``FieldState:361``,
``ArrayContainer:392``, ``reset:451``) and ``fdtdx/interfaces/state.py:8-19`` (``RecordingState``).
Arrays are ``torch.Tensor`` (CUDA) on the product path and ``numpy.ndarray`` when the oracle
drives the same containers; shapes and dtypes are the reference's (``initialization.py:598-842``).
"""

from __future__ import annotations

from dataclasses import dataclass, field, replace, fields as dc_fields
from typing import Any, Callable

import numpy as np

from fdtdx_b200.boundaries import (
    BaseBoundary,
    BlochBoundary,
    PerfectElectricConductor,
    PerfectlyMatchedLayer,
    PerfectMagneticConductor,
    SimulationObject,
    SimulationVolume,
)
from fdtdx_b200.detectors import Detector
from fdtdx_b200.sources import Source

PmlAuxField = dict  # name -> (psi_1, psi_2)
DetectorState = dict


def _map(x, fn: Callable):
    """tree-map over None / dict / tuple / list / array leaves."""
    if x is None:
        return None
    if isinstance(x, dict):
        return {k: _map(v, fn) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return type(x)(_map(v, fn) for v in x)
    if isinstance(x, (float, int)):
        return x
    return fn(x)


def _zeros_like(a):
    if isinstance(a, np.ndarray):
        return np.zeros_like(a)
    import torch

    return torch.zeros_like(a)


@dataclass
class RecordingState:
    data: dict
    state: dict = field(default_factory=dict)


@dataclass
class FieldState:
    E: Any
    H: Any
    psi_E: PmlAuxField
    psi_H: PmlAuxField
    dispersive_P_curr: Any = None
    dispersive_P_prev: Any = None


@dataclass
class ArrayContainer:
    fields: FieldState
    inv_permittivities: Any
    inv_permeabilities: Any
    detector_states: dict
    recording_state: RecordingState | None
    electric_conductivity: Any = None
    magnetic_conductivity: Any = None
    dispersive_c1: Any = None
    dispersive_c2: Any = None
    dispersive_c3: Any = None
    dispersive_c4: Any = None
    initial_inv_permittivities: Any = None

    # E, H convenience accessors used by a lot of reference user code (arrays.E / arrays.H)
    @property
    def E(self):
        return self.fields.E

    @property
    def H(self):
        return self.fields.H

    def aset(self, path: str, value: Any) -> "ArrayContainer":
        """Functional set with the reference's ``a->b`` path syntax (pytreeclass ``aset``)."""
        parts = path.split("->")
        if len(parts) == 1:
            return replace(self, **{parts[0]: value})
        if parts[0] != "fields" or len(parts) != 2:
            raise KeyError(path)
        return replace(self, fields=replace(self.fields, **{parts[1]: value}))

    def reset(self, reset_detector_states: bool = True, reset_recording_state: bool = False) -> "ArrayContainer":
        """``container.py:451-491``: zero E, H, psi, P (and detector states); keep materials."""
        new_fields = FieldState(**{f.name: _map(getattr(self.fields, f.name), _zeros_like) for f in dc_fields(FieldState)})
        det = self.detector_states
        if reset_detector_states:
            det = _map(det, _zeros_like)
        rec = self.recording_state
        if reset_recording_state and rec is not None:
            rec = RecordingState(data=_map(rec.data, _zeros_like), state=_map(rec.state, _zeros_like))
        return replace(self, fields=new_fields, detector_states=det, recording_state=rec)

    def map_arrays(self, fn: Callable) -> "ArrayContainer":
        kw = {}
        for f in dc_fields(ArrayContainer):
            v = getattr(self, f.name)
            if f.name == "fields":
                kw[f.name] = FieldState(**{g.name: _map(getattr(v, g.name), fn) for g in dc_fields(FieldState)})
            elif f.name == "recording_state":
                kw[f.name] = None if v is None else RecordingState(data=_map(v.data, fn), state=_map(v.state, fn))
            else:
                kw[f.name] = _map(v, fn)
        return ArrayContainer(**kw)

    def to_numpy(self) -> "ArrayContainer":
        def conv(a):
            if isinstance(a, np.ndarray):
                return a.copy()
            import torch

            t = a.detach().cpu()
            if t.dtype in (torch.bfloat16, torch.float8_e4m3fnuz, torch.float8_e4m3fn, torch.float8_e5m2, torch.float16):
                return _TorchLeaf(t)
            return t.numpy().copy()

        return self.map_arrays(conv)

    def to_torch(self, device="cuda") -> "ArrayContainer":
        import torch

        def conv(a):
            if isinstance(a, _TorchLeaf):
                return a.t.to(device)
            if isinstance(a, np.ndarray):
                return torch.from_numpy(np.ascontiguousarray(a)).to(device)
            return a.to(device)

        return self.map_arrays(conv)


class _TorchLeaf:
    """Wrapper for CPU tensors of dtypes numpy cannot hold (bf16 / fp8 recorder buffers)."""

    def __init__(self, t):
        self.t = t

    @property
    def shape(self):
        return tuple(self.t.shape)


SimulationState = tuple  # (time_step, ArrayContainer)


@dataclass
class ObjectContainer:
    """``container.py:33-355`` (list + the filters the step uses)."""

    object_list: list
    volume_idx: int = 0

    @property
    def volume(self) -> SimulationVolume:
        return self.object_list[self.volume_idx]

    @property
    def objects(self) -> list:
        return self.object_list

    @property
    def boundary_objects(self) -> list[BaseBoundary]:
        return [o for o in self.object_list if isinstance(o, BaseBoundary)]

    @property
    def pml_objects(self) -> list[PerfectlyMatchedLayer]:
        return [o for o in self.object_list if isinstance(o, PerfectlyMatchedLayer)]

    @property
    def pec_objects(self):
        return [o for o in self.object_list if isinstance(o, PerfectElectricConductor)]

    @property
    def pmc_objects(self):
        return [o for o in self.object_list if isinstance(o, PerfectMagneticConductor)]

    @property
    def periodic_objects(self):
        return [o for o in self.object_list if isinstance(o, BlochBoundary)]

    @property
    def sources(self) -> list[Source]:
        return [o for o in self.object_list if isinstance(o, Source)]

    @property
    def detectors(self) -> list[Detector]:
        return [o for o in self.object_list if isinstance(o, Detector)]

    @property
    def forward_detectors(self) -> list[Detector]:
        return [o for o in self.detectors if not o.inverse]

    @property
    def backward_detectors(self) -> list[Detector]:
        return [o for o in self.detectors if o.inverse]

    def __getitem__(self, name: str):
        for o in self.object_list:
            if o.name == name:
                return o
        raise KeyError(name)
