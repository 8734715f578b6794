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
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SLAB CHECK", "PASSED" if int(t.item()) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
