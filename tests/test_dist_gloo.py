"""world_size-2 gloo test of the x-slab decomposition and halo exchange.

The product stepper is CUDA-only, so here each rank drives the *oracle* on its slab extended by
the received halo plane (one ghost plane), through the same ``HaloExchange`` / ``slab_bounds`` host
logic the GPU runner uses.  The sharded run must equal the single-domain run - the equality test
the reference itself lacks."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fdtdx_b200 as fx
from fdtdx_b200.dist import HaloExchange, slab_bounds
from oracle import yee

F = np.float32
SHAPE = (12, 6, 8)
STEPS = 6


def _scene(shape, x_range=None):
    cfg = fx.SimulationConfig(time=5e-15, grid=fx.UniformGrid(spacing=50e-9))
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = [b for b in fx.boundary_objects_from_config(shape, cfg, "periodic") if b.axis != 0]  # x: zero halo
    rng = np.random.default_rng(7)
    inv_eps = (1 / (1 + rng.random((3, *SHAPE)))).astype(F)
    E = (1e-3 * rng.standard_normal((3, *SHAPE))).astype(F)
    H = (1e-3 * rng.standard_normal((3, *SHAPE))).astype(F)
    if x_range is not None:
        sl = (slice(None), slice(*x_range))
        inv_eps, E, H = inv_eps[sl].copy(), E[sl].copy(), H[sl].copy()
    objects, arrays, _, cfg, _ = fx.place_objects([vol, *bl], cfg, inv_permittivities=inv_eps)
    arrays.fields.E[...] = E
    arrays.fields.H[...] = H
    return objects, arrays, cfg


def _slab_objects(nx_local, cfg):
    shape = (nx_local, SHAPE[1], SHAPE[2])
    vol = fx.SimulationVolume(name="v", grid_slice_tuple=tuple((0, n) for n in shape))
    bl = [b for b in fx.boundary_objects_from_config(shape, cfg, "periodic") if b.axis != 0]
    return fx.ObjectContainer([vol, *bl])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x0, x1 = slab_bounds(SHAPE[0], world, rank)
    _, arrays, cfg = _scene((x1 - x0, SHAPE[1], SHAPE[2]), (x0, x1))
    hx = HaloExchange(rank, world, periodic_x=False)
    E, H, eps = arrays.fields.E, arrays.fields.H, arrays.inv_permittivities
    ny, nz = SHAPE[1], SHAPE[2]
    for t in range(STEPS):
        # E half-step: ghost plane x0-1 of H from the lower neighbour (zeros at the domain edge)
        send = torch.from_numpy(np.ascontiguousarray(H[1:3, -1]))
        recv = torch.zeros((2, ny, nz))
        hx.exchange(send, recv, None, None)
        ghost = np.zeros((3, 1, ny, nz), F)
        ghost[1:3, 0] = recv.numpy()
        ext = arrays.aset("fields->H", np.concatenate([ghost, H], axis=1)).aset("fields->E", np.concatenate([np.zeros_like(ghost), E], axis=1))
        ext = ext.aset("inv_permittivities", np.concatenate([eps[:, :1], eps], axis=1))
        E = yee.update_E(t, ext, _slab_objects(x1 - x0 + 1, cfg), cfg, True).fields.E[:, 1:]
        # H half-step: ghost plane x1 of E from the upper neighbour
        send = torch.from_numpy(np.ascontiguousarray(E[1:3, 0]))
        recv = torch.zeros((2, ny, nz))
        hx.exchange(None, None, send, recv)
        ghost = np.zeros((3, 1, ny, nz), F)
        ghost[1:3, 0] = recv.numpy()
        ext = arrays.aset("fields->E", np.concatenate([E, ghost], axis=1)).aset("fields->H", np.concatenate([H, np.zeros_like(ghost)], axis=1))
        ext = ext.aset("inv_permittivities", np.concatenate([eps, eps[:, -1:]], axis=1))
        H = yee.update_H(t, ext, _slab_objects(x1 - x0 + 1, cfg), cfg, True).fields.H[:, :-1]
    gathered = [torch.zeros((2, 3, x1 - x0, ny, nz)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(np.stack([E, H])))
    if rank == 0:
        full = torch.cat(gathered, dim=2).numpy()
        np.save(out, full)
    dist.destroy_process_group()


def test_two_rank_slab_run_equals_single_domain(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "sharded.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sharded = np.load(out)
    objects, arrays, cfg = _scene(SHAPE)
    st = (0, arrays)
    for _ in range(STEPS):
        st = yee.forward(st, cfg, objects, None, False, False, True)
    assert np.array_equal(sharded[0], st[1].fields.E), "sharded E differs from the single-domain run"
    assert np.array_equal(sharded[1], st[1].fields.H), "sharded H differs from the single-domain run"
