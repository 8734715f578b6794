"""Design devices and ``apply_params`` on the GPU.

Mirrors, for the part the inverse-design loop needs, ``fdtdx/objects/device/device.py:27-367`` (a
``Device`` = a box of design voxels with an ordered material list and a chain of parameter transforms),
``fdtdx/objects/device/parameters/{continuous,projection,discretization}.py`` (``GaussianSmoothing2D``,
``TanhProjection``, ``ClosestIndex`` with the straight-through estimator of ``core/jax/ste.py``) and
``fdtdx/fdtd/initialization.py:317-521`` (``apply_params``: latent parameters -> material indices on the
simulation grid -> ``inv_permittivities`` inside the device's grid slice; continuous devices interpolate
the permittivity linearly between their two materials, discrete devices select ``1 / eps`` per voxel).

Everything is torch ops on the arrays' device, out of place, so the gradient that
``reversible_fdtd`` / ``checkpointed_fdtd`` return for ``inv_permittivities`` flows on to the latent
parameters through ordinary autograd.  The voxel grid must tile the device's grid slice evenly (uniform
grids: the reference's physical-overlap resampling reduces to voxel repetition there).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from fdtdx_b200.boundaries import SimulationObject
from fdtdx_b200.initialization import Material

CONTINUOUS, BINARY, DISCRETE = "continuous", "binary", "discrete"


def straight_through_estimator(x, y):
    """``core/jax/ste.py:4-27``: forward value ``y``, gradient of ``x``."""
    return x - x.detach() + y.detach()


@dataclass
class GaussianSmoothing2D:
    """``continuous.py:172-267``: 2-D Gaussian blur of the (nx, ny, 1)-shaped design (kernel
    ``6 std + 1`` wide, normalised, edge-repeat padding)."""

    std_discrete: int = 1
    output_type = None  # same as input

    def __call__(self, x, **kwargs):
        import torch
        import torch.nn.functional as F

        vert = list(x.shape).index(1)
        x2 = x.squeeze(vert)
        size = 6 * self.std_discrete + 1
        c = torch.arange(-(size // 2), size // 2 + 1, dtype=x.dtype, device=x.device)
        k = torch.exp(-(c[:, None] ** 2 + c[None, :] ** 2) / (2 * self.std_discrete**2))
        k = k / k.sum()
        pad = size // 2
        xp = F.pad(x2[None, None], (pad, pad, pad, pad), mode="replicate")
        return F.conv2d(xp, k[None, None])[0, 0].unsqueeze(vert)


@dataclass
class TanhProjection:
    """``projection.py:11-46, 182-217``; needs ``beta=`` among the ``apply_params`` keyword arguments."""

    projection_midpoint: float = 0.5
    output_type = None

    def __call__(self, x, **kwargs):
        import torch

        if "beta" not in kwargs:
            raise Exception("TanhProjection needs the beta parameter as additional keyword argument!")
        beta, eta = float(kwargs["beta"]), self.projection_midpoint
        if beta == 0:
            return x.clamp(0, 1)
        if math.isinf(beta):
            return (x > eta).to(x.dtype)
        return (math.tanh(beta * eta) + torch.tanh(beta * (x - eta))) / (math.tanh(beta * eta) + math.tanh(beta * (1 - eta)))


@dataclass
class ClosestIndex:
    """``discretization.py:21-90`` (``mapping_from_inverse_permittivities=False``): round to the nearest
    material index, gradient straight through."""

    output_type = DISCRETE
    _num_materials: int = 2

    def __call__(self, x, **kwargs):
        import torch

        return straight_through_estimator(x, torch.clamp(torch.round(x), 0, self._num_materials - 1))


@dataclass
class Device(SimulationObject):
    """A design region: ``materials`` (name -> Material, ordered by permittivity like
    ``compute_ordered_material_name_tuples``), a voxel grid that tiles ``grid_slice_tuple`` and a chain of
    parameter transforms.  The output type is that of the last transform (continuous if none sets one)."""

    materials: dict = field(default_factory=dict)
    voxel_grid_shape: tuple[int, int, int] | None = None  # default: one voxel per grid cell
    param_transforms: Sequence[object] = ()

    def __post_init__(self):
        if len(self.materials) < 2:
            raise Exception(f"Invalid materials (need two or more): {self.materials}")
        for t in self.param_transforms:
            if isinstance(t, ClosestIndex):
                t._num_materials = len(self.materials)

    @property
    def matrix_voxel_grid_shape(self) -> tuple[int, int, int]:
        return tuple(self.voxel_grid_shape) if self.voxel_grid_shape is not None else tuple(self.grid_shape)

    @property
    def output_type(self) -> str:
        out = CONTINUOUS
        for t in self.param_transforms:
            if getattr(t, "output_type", None) is not None:
                out = t.output_type
        return out

    def ordered_materials(self):
        def first(v):
            return float(np.atleast_1d(np.asarray(v, np.float64)).reshape(-1)[0])

        return sorted(self.materials.items(), key=lambda m: (first(m[1].permittivity), first(m[1].permeability), first(m[1].electric_conductivity), first(m[1].magnetic_conductivity)))

    def init_params(self, seed: int = 0, device="cpu"):
        """Uniform [0, 1) latent parameters on the voxel grid (``device.py:317-340``)."""
        import torch

        g = torch.Generator(device="cpu").manual_seed(seed)
        return torch.rand(self.matrix_voxel_grid_shape, generator=g, dtype=torch.float32).to(device)

    def __call__(self, params, expand_to_sim_grid: bool = False, **transform_kwargs):
        """Latent parameters -> material indices (``device.py:342-367``)."""
        x = params
        for t in self.param_transforms:
            x = t(x, **transform_kwargs)
        if expand_to_sim_grid:
            for axis, (n_vox, n_grid) in enumerate(zip(self.matrix_voxel_grid_shape, self.grid_shape)):
                if n_grid % n_vox != 0:
                    raise NotImplementedError(f"voxel grid {self.matrix_voxel_grid_shape} must tile the device's grid slice {self.grid_shape} evenly")
                if n_grid != n_vox:
                    x = x.repeat_interleave(n_grid // n_vox, dim=axis)
        return x


def apply_params(arrays, objects, params, key=None, **transform_kwargs):
    """``initialization.py:317-521``.  ``params``: ``{device.name: latent tensor}`` on the arrays' device.
    Returns ``(arrays, objects, info)``; ``arrays.inv_permittivities`` is a new tensor that depends
    differentiably on the parameters.  Sources / detectors whose mode overlaps a device are not re-solved
    here (place them on fixed cross-sections, as the shipped examples do)."""
    import torch

    devices = [o for o in objects.object_list if isinstance(o, Device)]
    if not devices:
        if params:
            raise Exception("apply_params: parameters given but the scene has no Device")
        return arrays, objects, {}
    inv_eps = arrays.inv_permittivities
    if not torch.is_tensor(inv_eps):
        raise RuntimeError("apply_params runs on device tensors: move the container with arrays.to_torch('cuda') first")
    ncomp = int(inv_eps.shape[0])
    if ncomp == 9:
        raise NotImplementedError("apply_params for fully anisotropic devices")
    initial = getattr(arrays, "initial_inv_permittivities", None)
    if initial is not None:
        inv_eps = initial
    for dev in devices:
        idx = dev(params[dev.name], expand_to_sim_grid=True, **transform_kwargs)
        perms = []
        for _, m in dev.ordered_materials():
            p = np.atleast_1d(np.asarray(m.permittivity, np.float64)).reshape(-1)
            if p.size == 9:
                p = p[[0, 4, 8]]
            if p.size == 1 and ncomp == 3:
                p = np.repeat(p, 3)
            perms.append(p[:ncomp])
        allowed = torch.as_tensor(np.stack(perms), dtype=inv_eps.dtype, device=inv_eps.device)  # (n_materials, ncomp)
        if dev.output_type == CONTINUOUS:
            # linear interpolation between the two device materials in permittivity, then inverted
            perm = allowed[0][:, None, None, None] + idx[None] * (allowed[1] - allowed[0])[:, None, None, None]
            new_slice = 1.0 / perm
        else:
            inv_allowed = 1.0 / allowed
            vals = inv_allowed[idx.detach().to(torch.long)].movedim(-1, 0)  # (ncomp, *grid)
            new_slice = straight_through_estimator(idx[None].expand_as(vals), vals)
        sl = (slice(None), *dev.grid_slice)
        pad = [(lo, n - hi) for (lo, hi), n in zip(dev.grid_slice_tuple, inv_eps.shape[1:])]
        # out-of-place write (jnp .at[].set): the result depends on the parameters inside the device box
        # and on the previous array outside it
        mask = torch.zeros(inv_eps.shape[1:], dtype=torch.bool, device=inv_eps.device)
        mask[dev.grid_slice] = True
        full = torch.nn.functional.pad(new_slice, (pad[2][0], pad[2][1], pad[1][0], pad[1][1], pad[0][0], pad[0][1]))
        inv_eps = torch.where(mask[None], full, inv_eps)
        del sl
    arrays = arrays.aset("inv_permittivities", inv_eps.contiguous())
    return arrays, objects, {}
