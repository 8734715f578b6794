"""torchrun --nproc-per-node N scripts/check_slabs.py : x-slab sharded run == single-GPU run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import fdtdx_b200 as fx
from fdtdx_b200 import workloads as W
from fdtdx_b200.dist import SlabRunner, slab_bounds
from fdtdx_b200.fdtd import get_plan

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
ok = True
for name, build, steps in (
    ("box", lambda xr: W.build_box((32 * world, 40, 64), device=dev, x_range=xr, thickness=6), 40),
    ("box_flat_tiles", lambda xr: W.build_box((24 * world, 45, 72), device=dev, x_range=xr, thickness=6), 40),  # 18-quad rows: flat tiles, 14 rows per CTA
    ("coupler", lambda xr: W.build_coupler(5, device=dev, n_copies=world, x_range=xr), 60),
):
    objects, arrays_full, cfg = build(None)
    nx = objects.volume.grid_shape[0]
    for overlap, halo in ((False, "nccl"), (True, "nccl"), (False, "peer")):
        x0, x1 = slab_bounds(nx, world, rank)
        objects_s, arrays, cfg_s = build((x0, x1))
        runner = SlabRunner(objects_s, cfg_s, arrays, (x0, x1), rank, world, overlap=overlap, halo=halo)
        runner.run(0, steps, record_detectors=True)
        torch.cuda.synchronize()
        dist.barrier()  # peer mode: neighbours read this rank's arrays in place until they are done too
        # single-GPU reference on every rank
        objects_f, arrays_f, cfg_f = build(None)
        plan = get_plan(arrays_f, objects_f, cfg_f)
        plan.run_forward(0, steps, True, False, True)
        torch.cuda.synchronize()
        dE = float((arrays.fields.E - arrays_f.fields.E[:, x0:x1]).abs().max())
        dH = float((arrays.fields.H - arrays_f.fields.H[:, x0:x1]).abs().max())
        mx = float(arrays_f.fields.E.abs().max())
        ddet = 0.0
        for k, v in arrays.detector_states.items():
            for k2, v2 in v.items():
                ddet = max(ddet, float((v2 - arrays_f.detector_states[k][k2]).abs().max()))
        good = dE == 0.0 and dH == 0.0 and ddet == 0.0 and mx > 0
        ok = ok and good
        print(f"[rank {rank}] {name} halo={halo} overlap={overlap}: max|dE|={dE:.3e} max|dH|={dH:.3e} det={ddet:.3e} (|E|max={mx:.3e}) {'OK' if good else 'MISMATCH'}", flush=True)
        objects_f.__dict__.pop("_plan_cache", None)
# ---- recorder + time-reversed pass on slabs (peer-memory halo): forward with PML-interface recording,
# full_backward with interface replay, PML reset and an inverse detector, vs the single-GPU run
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from scenes import build_scene
from fdtdx_b200.dist import shard_arrays

for rec_name, modules in (("f32", []), ("every3+bf16", [fx.LinearReconstructEveryK(k=3), fx.DtypeConversion(dtype="bfloat16")])):
    nxs = 16 * world
    objects, arrays_np, cfg = build_scene(shape=(nxs, 14, 20), thickness=4, source="plane_z", time=5e-15, recorder=fx.Recorder(modules=modules), eps_tier=3, sigma_E=True)
    dets = [fx.EnergyDetector(name=f"inv{r}", grid_slice_tuple=((16 * r + 2, 16 * r + 13), (1, 13), (2, 18)), as_slices=True, inverse=True) for r in range(world)]
    dets += [fx.FieldDetector(name=f"fld{r}", grid_slice_tuple=((16 * r + 1, 16 * r + 15), (2, 12), (3, 17)), switch=fx.OnOffSwitch(interval=2)) for r in range(world)]
    objects, arrays_np, _, cfg, _ = fx.place_objects(list(objects.object_list) + dets, cfg, inv_permittivities=arrays_np.inv_permittivities,
                                                     electric_conductivity=arrays_np.electric_conductivity)
    T = cfg.time_steps_total
    x0, x1 = slab_bounds(nxs, world, rank)
    arrays = shard_arrays(arrays_np, objects, (x0, x1), dev)
    runner = SlabRunner(objects, cfg, arrays, (x0, x1), rank, world, halo="peer")
    good = runner.peer
    if good:
        # forward and time-reversed pass are enqueued back to back, with no host synchronisation in between
        runner.run(0, T, record_detectors=True, record_boundaries=True)
        fwd = {"E": arrays.fields.E.clone(), "H": arrays.fields.H.clone(), "rec": {k: v.clone() for k, v in arrays.recording_state.data.items()}}
        runner.run_reverse(T, T, record_detectors=True, reset_fields=True)
        torch.cuda.synchronize()
        dist.barrier()
        full = arrays_np.to_torch(dev)
        st = fx.run_fdtd(full, objects, cfg)
        torch.cuda.synchronize()
        fE = float((fwd["E"] - st[1].fields.E[:, x0:x1]).abs().max())
        fH = float((fwd["H"] - st[1].fields.H[:, x0:x1]).abs().max())
        frec = 0.0
        for k, v in fwd["rec"].items():
            r = st[1].recording_state.data[k]
            r = r if k.split("_")[2] == "x" else r[:, :, x0:x1]
            frec = max(frec, float((v.float() - r.float()).abs().max()))
        print(f"[rank {rank}] recorder({rec_name}) forward on slabs: max|dE|={fE:.3e} max|dH|={fH:.3e} recorder buffers {frec:.3e}", flush=True)
        good = good and fE == 0.0 and fH == 0.0 and frec == 0.0
        st = fx.full_backward(st, objects, cfg, record_detectors=True, reset_fields=True)
        torch.cuda.synchronize()
        ref = st[1]
        dE = float((arrays.fields.E - ref.fields.E[:, x0:x1]).abs().max())
        dH = float((arrays.fields.H - ref.fields.H[:, x0:x1]).abs().max())
        ddet, live = 0.0, 0.0
        for k, v in arrays.detector_states.items():
            for k2, v2 in v.items():
                ddet = max(ddet, float((v2 - ref.detector_states[k][k2]).abs().max()))
                live = max(live, float(ref.detector_states[k][k2].abs().max()))
        good = good and dE == 0.0 and dH == 0.0 and ddet == 0.0 and live > 0 and int(runner.plan.lib.fdtdx_b200_peer_status(runner.plan.h)) == 0
        print(f"[rank {rank}] recorder({rec_name}) + full_backward on slabs: max|dE|={dE:.3e} max|dH|={dH:.3e} det={ddet:.3e} (|det|max={live:.3e}) {'OK' if good else 'MISMATCH'}", flush=True)
        objects.__dict__.pop("_plan_cache", None)
    else:
        print(f"[rank {rank}] recorder + full_backward on slabs: peer-memory halo unavailable, skipped", flush=True)
        good = os.environ.get("FDTDX_B200_PEER_FAIL") is not None
    ok = ok and good
# ---- detectors straddling slab edges: every rank samples its part of the region, the
# co-location stencil at a slab's first plane reads the lower neighbour's last plane (exchanged on the
# steps the detector is on), and one merge after the run rebuilds the whole-region states
from fdtdx_b200.dist import merge_detector_states  # noqa: E402,F401

nxs = 16 * world
for halo, overlap in (("peer", False), ("nccl", True)):
    objects, arrays_np, cfg = build_scene(shape=(nxs, 14, 36), thickness=4, source="plane_z", time=5e-15, eps_tier=3)
    wc = fx.WaveCharacter(wavelength=0.8e-6)
    dets = [
        fx.EnergyDetector(name="video", grid_slice_tuple=((0, nxs), (0, 14), (0, 36)), as_slices=True, switch=fx.OnOffSwitch(interval=3)),   # row-marching kernels, YZ mean over ranks
        fx.EnergyDetector(name="cube", grid_slice_tuple=((3, nxs - 2), (2, 12), (1, 35))),                                                  # row-marching, region-shaped
        fx.FieldDetector(name="block", grid_slice_tuple=((5, nxs - 5), (3, 11), (4, 20)), switch=fx.OnOffSwitch(interval=2)),               # generic kernels
        fx.FieldDetector(name="fmean", grid_slice_tuple=((2, nxs - 3), (2, 12), (5, 30)), reduce_volume=True, components=("Ex", "Hz")),       # weighted mean, global weight sum
        fx.EnergyDetector(name="etot", grid_slice_tuple=((0, nxs), (0, 14), (0, 36)), reduce_volume=True),
        fx.PoyntingFluxDetector(name="flux", grid_slice_tuple=((0, nxs), (0, 14), (28, 29)), direction="+"),                                # z-normal plane across all ranks
        fx.PhasorDetector(name="ph", grid_slice_tuple=((1, nxs - 1), (7, 8), (2, 34)), wave_characters=(wc,)),                              # y-normal plane across all ranks
        fx.EnergyDetector(name="fixed", grid_slice_tuple=((0, nxs), (0, 14), (0, 36)), as_slices=True, x_slice=(nxs // 2 + 3) * 50e-9, y_slice=3 * 50e-9, z_slice=9 * 50e-9),
    ]
    objects, arrays_np, _, cfg, _ = fx.place_objects(list(objects.object_list) + dets, cfg, inv_permittivities=arrays_np.inv_permittivities)
    T = cfg.time_steps_total
    x0, x1 = slab_bounds(nxs, world, rank)
    arrays = shard_arrays(arrays_np, objects, (x0, x1), dev, config=cfg)
    runner = SlabRunner(objects, cfg, arrays, (x0, x1), rank, world, overlap=overlap, halo=halo)
    runner.run(0, T, record_detectors=True)
    torch.cuda.synchronize()
    dist.barrier()
    merged = runner.gather_detector_states()
    st = fx.run_fdtd(arrays_np.to_torch(dev), objects, cfg)
    torch.cuda.synchronize()
    dE = float((arrays.fields.E - st[1].fields.E[:, x0:x1]).abs().max())
    good = dE == 0.0
    for d, stt in st[1].detector_states.items():
        for k, v in stt.items():
            ref = v.cpu().numpy()
            got = merged[d][k]
            exact = d in ("cube", "block", "ph", "fixed") or (d == "video" and k != "YZ Plane")
            err = float(np.abs(got - ref).max()) / max(float(np.abs(ref).max()), 1e-30)
            fine = (err == 0.0) if exact else (err <= 2e-6)
            good = good and fine and got.shape == ref.shape and float(np.abs(ref).max()) > 0
            if rank == 0:
                print(f"  straddling {d}/{k} halo={halo}: max rel err {err:.2e} {'(bit-exact)' if exact else ''} {'OK' if fine else 'MISMATCH'}", flush=True)
    print(f"[rank {rank}] straddling detectors halo={halo}{'' if runner.peer or halo != 'peer' else ' (fell back to nccl)'}: fields max|dE|={dE:.3e} {'OK' if good else 'MISMATCH'}", flush=True)
    objects.__dict__.pop("_plan_cache", None)
    ok = ok and good
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SLAB CHECK", "PASSED" if int(t.item()) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
