"""``apply_params`` on device tensors (fdtd/initialization.py:317-521,
objects/device/*): latent parameters -> inv_permittivities -> reversible FDTD -> loss, and the
gradient all the way back to the parameters."""

import numpy as np
import pytest
import torch

import fdtdx_b200 as fx
from fdtdx_b200.device import apply_params as apply_params_torch
from oracle import yee_torch
from scenes import build_scene, rel_l2


def _device(transforms, voxels=None):
    return fx.Device(name="design", grid_slice_tuple=((3, 9), (2, 8), (4, 10)), voxel_grid_shape=voxels,
                     materials={"air": fx.Material(permittivity=1.0), "si": fx.Material(permittivity=12.25)}, param_transforms=transforms)


def _scene(dev):
    objects, arrays, cfg = build_scene(boundaries="periodic", source="plane_z", detectors=("field", "phasor"), time=3e-15, recorder=fx.Recorder(modules=[]))
    objs = list(objects.object_list) + [dev]
    objects, arrays, _, cfg, _ = fx.place_objects(objs, cfg, inv_permittivities=arrays.inv_permittivities)
    return objects, arrays, cfg


def test_continuous_and_discrete_mappings_on_cpu_tensors():
    """Known answers of the parameter mapping itself (no FDTD): voxel repetition, the linear
    permittivity interpolation, nearest-index selection with a straight-through gradient."""
    dev = _device([], voxels=(3, 3, 1))
    objects, arrays, cfg = _scene(dev)
    a = arrays.to_torch("cpu")
    p = torch.linspace(0.0, 1.0, 9).reshape(3, 3, 1).requires_grad_(True)
    out, _, _ = apply_params_torch(a, objects, {"design": p})
    blk = out.inv_permittivities[0, 3:9, 2:8, 4:10]
    want = 1.0 / (1.0 + p.detach().repeat_interleave(2, 0).repeat_interleave(2, 1).repeat_interleave(6, 2) * 11.25)
    assert torch.allclose(blk, want.to(blk.dtype), rtol=1e-6)
    assert torch.equal(out.inv_permittivities[0, :3], a.inv_permittivities[0, :3])  # untouched outside the box
    blk.sum().backward()
    assert p.grad is not None and float(p.grad.abs().min()) > 0
    # discrete: 0.3 -> air, 0.8 -> silicon; d out / d p = 1 (straight through)
    dev = _device([fx.ClosestIndex()], voxels=(1, 1, 1))
    objects, arrays, cfg = _scene(dev)
    for val, eps in ((0.3, 1.0), (0.8, 12.25)):
        p = torch.full((1, 1, 1), val, requires_grad=True)
        out, _, _ = apply_params_torch(arrays.to_torch("cpu"), objects, {"design": p})
        blk = out.inv_permittivities[0, 3:9, 2:8, 4:10]
        assert torch.allclose(blk, torch.full_like(blk, 1.0 / eps))
        blk.sum().backward()
        assert float(p.grad) == pytest.approx(blk.numel())
    # smoothing keeps a constant design constant (normalised kernel, edge-repeat padding); tanh projection
    # maps the midpoint to 0.5 and is monotone
    g = fx.GaussianSmoothing2D(std_discrete=2)
    assert torch.allclose(g(torch.full((7, 5, 1), 0.37)), torch.full((7, 5, 1), 0.37), atol=1e-6)
    tp = fx.TanhProjection()
    x = torch.tensor([0.1, 0.5, 0.9])
    y = tp(x, beta=8.0)
    assert float(y[1]) == pytest.approx(0.5, abs=1e-6) and float(y[0]) < 0.1 and float(y[2]) > 0.9


@pytest.mark.gpu
def test_gradient_reaches_the_latent_parameters():
    """params -> (smoothing, tanh projection) -> inv_eps -> reversible_fdtd -> loss.backward(): d loss / d
    params from the CUDA adjoint equals float64 autograd through the same mapping + the restated forward run."""
    dev = _device([fx.GaussianSmoothing2D(std_discrete=1), fx.TanhProjection()], voxels=(6, 6, 1))
    objects, arrays, cfg = _scene(dev)
    T = cfg.time_steps_total
    p0 = dev.init_params(seed=3)

    def loss_of(det, E):
        return (det["phasor"]["phasor"].abs() ** 2).sum() * 3.0 + (det["field"]["fields"] ** 2).sum() + (E * E).sum() * 0.1

    # reference: float64 on the CPU
    p_ref = p0.clone().double().requires_grad_(True)
    a64 = arrays.to_torch("cpu")
    a64 = a64.aset("inv_permittivities", a64.inv_permittivities.double())
    out64, _, _ = apply_params_torch(a64, objects, {"design": p_ref}, beta=4.0)
    E, H, det = yee_torch.run_forward(arrays.reset(), objects, cfg, T, inv_eps=out64.inv_permittivities, dtype=torch.float64)
    loss_ref = loss_of(det, E)
    loss_ref.backward()
    # product: CUDA
    p = p0.clone().cuda().requires_grad_(True)
    dev_arrays, _, _ = fx.apply_params(arrays.to_torch("cuda"), objects, {"design": p}, beta=4.0)
    assert dev_arrays.inv_permittivities.requires_grad
    t_end, out = fx.run_fdtd(dev_arrays, objects, cfg)
    loss = loss_of(out.detector_states, out.fields.E)
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    loss.backward()
    assert float(p_ref.grad.abs().max()) > 0
    assert rel_l2(p.grad.cpu().numpy(), p_ref.grad.numpy()) <= 1e-4
