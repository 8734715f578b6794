"""Unfold helpers for symmetry-reduced runs (reference: ``fdtd/symmetry.py:311-753`` and the parity / index-map tables of
``core/physics/symmetry.py:18-107, 184-240``): pure post-processing on the arrays a reduced run returns - the time loop's
own symmetry rules (one-sided wrap, detector mirror) live in the kernels.

Arrays may be ``torch`` tensors (any device) or NumPy arrays; the result is of the same kind.

A host-mirror object carries no placement history, so "the symmetry plane clipped this detector" (the reference's
``straddles_symmetry_plane``: ``unreduced_grid_slice_tuple[axis][0] < 0``) is read from that attribute when the object
has it and otherwise taken to be "its slice starts at index 0 of a symmetric axis".
"""

from __future__ import annotations

from typing import Literal

import numpy as np

_COMPONENT_NAMES = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
_COMPONENT_SPEC = (("E", 0), ("E", 1), ("E", 2), ("H", 0), ("H", 1), ("H", 2))


# ------------------------------------------------------------------------------------------ array-kind shims
def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _flip(a, axis: int):
    if _is_torch(a):
        return a.flip(axis)
    return np.flip(a, axis=axis)


def _cat(parts, axis: int):
    if _is_torch(parts[0]):
        import torch

        return torch.cat(list(parts), dim=axis)
    return np.concatenate(list(parts), axis=axis)


def _slice_axis(a, axis: int, start: int, stop: int | None = None):
    index = [slice(None)] * a.ndim
    index[axis] = slice(start, stop)
    return a[tuple(index)]


def _factor_like(values, shape, like):
    if _is_torch(like):
        import torch

        return torch.tensor(values, dtype=like.real.dtype if like.is_complex() else like.dtype, device=like.device).reshape(shape)
    return np.asarray(values, dtype=like.real.dtype).reshape(shape)


# ------------------------------------------------------------------------------------------ parity / index-map tables
def field_component_parity(field_type: Literal["E", "H"], component: int, axis: int, wall: int) -> int:
    """+1 (even) / -1 (odd) of a field component under reflection about a mirror plane normal to ``axis``
    (``core/physics/symmetry.py:18-49``); ``wall``: -1 electric (PEC), +1 magnetic (PMC)."""
    normal = component == axis
    if wall == -1:
        return (1 if normal else -1) if field_type == "E" else (-1 if normal else 1)
    if wall == 1:
        return (-1 if normal else 1) if field_type == "E" else (1 if normal else -1)
    raise ValueError(f"wall must be -1 (PEC) or +1 (PMC), got {wall}")


def component_sits_on_plane(field_type: Literal["E", "H"], component: int, axis: int) -> bool:
    """Whether the component's samples along ``axis`` include the plane itself (``symmetry.py:52-73``)."""
    if field_type == "E":
        return component != axis
    if field_type == "H":
        return component == axis
    raise ValueError(f"field_type must be 'E' or 'H', got {field_type!r}")


def mirror_pairs_on_plane(field_type: Literal["E", "H"], component: int, axis: int, wall: int) -> bool:
    """Samples pair as m +- j about a shared row (electric plane, on-plane component) instead of a plain flip
    (``symmetry.py:76-107``)."""
    return wall == -1 and component_sits_on_plane(field_type, component, axis)


def mirror_extend_low_side(array, axis: int, parity: int, on_plane: bool):
    """The mirrored (discarded) half in ascending order, so that ``cat([low, array])`` is the full array
    (``symmetry.py:217-240``): a plain flip, or - samples on the plane - index 0 is its own mirror, indices 1..n-1
    produce images and the outermost missing sample repeats its neighbour."""
    if not on_plane:
        return parity * _flip(array, axis)
    mirrored = parity * _flip(_slice_axis(array, axis, 1), axis)
    return _cat([_slice_axis(mirrored, axis, 0, 1), mirrored], axis)


def _check_has_symmetry(symmetry) -> None:
    if not any(s != 0 for s in symmetry):
        raise ValueError(
            "Nothing to unfold: this simulation has no symmetry (config.symmetry is (0, 0, 0)). "
            "The unfold helpers are only meaningful for symmetry-reduced simulations."
        )


# ------------------------------------------------------------------------------------------ fields / generic arrays
def unfold_fields(field, symmetry, field_type: Literal["E", "H"]):
    """Full-domain ``(3, Nx, Ny, Nz)`` field from the reduced one (``fdtd/symmetry.py:319-366``): each symmetric axis is
    doubled, the mirror image (per-component parity and index map) in front of the kept half."""
    if field_type not in ("E", "H"):
        raise ValueError(f"field_type must be 'E' or 'H', got {field_type!r}")
    _check_has_symmetry(symmetry)
    arr = field
    for a in range(3):
        if symmetry[a] == 0:
            continue
        comps = []
        for c in range(3):
            single = arr[c : c + 1]
            low = mirror_extend_low_side(single, a + 1, field_component_parity(field_type, c, a, symmetry[a]),
                                         mirror_pairs_on_plane(field_type, c, a, symmetry[a]))
            comps.append(_cat([low, single], a + 1))
        arr = _cat(comps, 0)
    return arr


def unfold_array(arr, symmetry, spatial_axes, signs=None, on_plane_axes=()):
    """Mirror-and-concatenate along each symmetric axis (``fdtd/symmetry.py:434-474``); ``signs`` maps a physical axis
    to a broadcastable sign applied to the mirror image."""
    _check_has_symmetry(symmetry)
    for a in range(3):
        if symmetry[a] == 0:
            continue
        ax = spatial_axes[a]
        sign = 1.0 if signs is None else signs.get(a, 1.0)
        if a in on_plane_axes:
            mirror = mirror_extend_low_side(arr, ax, 1, True) * sign
        else:
            mirror = _flip(arr, ax) * sign
        arr = _cat([mirror, arr], ax)
    return arr


# ------------------------------------------------------------------------------------------ detector states
def straddles_symmetry_plane(obj, axis: int) -> bool:
    un = getattr(obj, "unreduced_grid_slice_tuple", None)
    if un is not None:
        return un[axis][0] < 0
    return obj.grid_slice_tuple[axis][0] == 0


def _colocated_on_plane_axes(detector, touched):
    """Co-located (exact-interpolation) samples sit at integer x, y: on an ELECTRIC plane normal to x or y they pair as
    m +- j; raw staggered components fall back to the plain flip (``fdtd/symmetry.py:477-496``)."""
    if not getattr(detector, "exact_interpolation", False):
        return ()
    return tuple(a for a in (0, 1) if touched[a] == -1)


def _stored_component_spec(components):
    return [_COMPONENT_SPEC[i] for i, name in enumerate(_COMPONENT_NAMES) if name in components]


def _component_signs(spec, touched, like, component_axis: int):
    out = {}
    for a in range(3):
        if touched[a] == 0:
            continue
        vals = [float(field_component_parity(ft, ca, a, touched[a])) for ft, ca in spec]
        shape = [1] * like.ndim
        shape[component_axis] = len(vals)
        out[a] = _factor_like(vals, shape, like)
    return out


def _reduce_factor(parities, like, component_axis: int, mean: bool):
    """prod over touched axes of (1 + parity) for a sum, (1 + parity) / 2 for a mean (``fdtd/symmetry.py:522-551``)."""
    factors = []
    for per_axis in parities:
        f = 1.0
        for p in per_axis:
            f *= (1 + p) / 2 if mean else (1 + p)
        factors.append(f)
    shape = [1] * like.ndim
    shape[component_axis] = len(factors)
    return _factor_like(factors, shape, like)


def _poynting_parity(component: int, axis: int, wall: int) -> int:
    j, k = (x for x in range(3) if x != component)
    return field_component_parity("E", j, axis, wall) * field_component_parity("H", k, axis, wall)


def _unfold_poynting(detector, state, touched):
    arr = state["poynting_flux"]
    on_plane = _colocated_on_plane_axes(detector, touched)
    keep_all = bool(getattr(detector, "keep_all_components", False))
    components = (0, 1, 2) if keep_all else (detector.propagation_axis,)
    touched_axes = [a for a in range(3) if touched[a] != 0]
    if detector.reduce_volume:
        parities = [[_poynting_parity(i, a, touched[a]) for a in touched_axes] for i in components]
        if keep_all:
            return {"poynting_flux": arr * _reduce_factor(parities, arr, 1, mean=False)}
        scalar = 1.0
        for p in parities[0]:
            scalar *= 1 + p
        return {"poynting_flux": arr * scalar}
    if keep_all:  # (T, 3, nx, ny, nz)
        signs = {a: _factor_like([float(_poynting_parity(i, a, touched[a])) for i in components], (1, 3, 1, 1, 1), arr) for a in touched_axes}
        return {"poynting_flux": unfold_array(arr, touched, (2, 3, 4), signs, on_plane)}
    p = detector.propagation_axis
    signs = {a: float(_poynting_parity(p, a, touched[a])) for a in touched_axes}
    return {"poynting_flux": unfold_array(arr, touched, (1, 2, 3), signs, on_plane)}


def _unfold_energy_slices(state, touched, on_plane_axes=()):
    planes = {"XY Plane": (0, 1), "XZ Plane": (0, 2), "YZ Plane": (1, 2)}
    out = {}
    for key, phys_axes in planes.items():
        arr = state[key]
        sub, spatial_axes = [0, 0, 0], [0, 0, 0]
        for arr_axis, phys in enumerate(phys_axes, start=1):
            sub[phys] = touched[phys]
            spatial_axes[phys] = arr_axis
        if any(s != 0 for s in sub):
            out[key] = unfold_array(arr, tuple(sub), tuple(spatial_axes), on_plane_axes=tuple(a for a in on_plane_axes if a in phys_axes))
        else:
            out[key] = arr
    return out


def _unfold_one_detector(detector, state, touched, count: int):
    from fdtdx_b200 import detectors as D

    if isinstance(detector, D.PhasorDetector):  # incl. the mode-overlap / phasor-flux detectors built on it
        spec = _stored_component_spec(detector.components)
        arr = state["phasor"]
        if detector.reduce_volume:  # (1, nf, ncomp): mean reduction
            parities = [[field_component_parity(ft, ca, a, touched[a]) for a in range(3) if touched[a]] for ft, ca in spec]
            return {"phasor": arr * _reduce_factor(parities, arr, 2, mean=True)}
        signs = _component_signs(spec, touched, arr, 2)  # (1, nf, ncomp, nx, ny, nz)
        return {"phasor": unfold_array(arr, touched, (3, 4, 5), signs, _colocated_on_plane_axes(detector, touched))}
    if isinstance(detector, D.FieldDetector):
        spec = _stored_component_spec(detector.components)
        arr = state["fields"]
        if detector.reduce_volume:  # (T, ncomp): mean reduction
            parities = [[field_component_parity(ft, ca, a, touched[a]) for a in range(3) if touched[a]] for ft, ca in spec]
            return {"fields": arr * _reduce_factor(parities, arr, 1, mean=True)}
        signs = _component_signs(spec, touched, arr, 1)  # (T, ncomp, nx, ny, nz)
        return {"fields": unfold_array(arr, touched, (2, 3, 4), signs, _colocated_on_plane_axes(detector, touched))}
    if isinstance(detector, D.EnergyDetector):
        if detector.as_slices:
            return _unfold_energy_slices(state, touched, _colocated_on_plane_axes(detector, touched))
        if detector.reduce_volume:  # energy density is even: x2 per plane
            return {"energy": state["energy"] * (2**count)}
        return {"energy": unfold_array(state["energy"], touched, (1, 2, 3), on_plane_axes=_colocated_on_plane_axes(detector, touched))}
    if type(detector).__name__ == "PoyntingFluxDetector":
        return _unfold_poynting(detector, state, touched)
    raise NotImplementedError(
        f"unfold_detector_states does not know how to unfold detector type {type(detector).__name__!r}. "
        "Unfold the fields with unfold_fields instead."
    )


def unfold_detector_states(arrays, objects, config):
    """Full-domain detector states from the reduced run's (``fdtd/symmetry.py:699-753``): spatial outputs are mirrored per
    component, ``reduce_volume`` sums / means rescaled per component, energy slices mirrored along their in-plane
    symmetric axes; detectors that no symmetry plane clipped are returned unchanged."""
    _check_has_symmetry(config.symmetry)
    by_name = {d.name: d for d in objects.detectors}
    new_states = {}
    for name, state in arrays.detector_states.items():
        det = by_name.get(name)
        if det is None:
            new_states[name] = state
            continue
        touched = tuple(config.symmetry[a] if straddles_symmetry_plane(det, a) else 0 for a in range(3))
        count = sum(1 for a in range(3) if touched[a] != 0)
        new_states[name] = state if count == 0 else _unfold_one_detector(det, state, touched, count)
    return arrays.aset("detector_states", new_states)
