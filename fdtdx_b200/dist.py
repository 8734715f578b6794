"""x-slab multi-GPU runner: one process per GPU, one halo plane per half-step.

The reference shards every ``(., Nx, Ny, Nz)`` array on x with a ``NamedSharding`` and lets XLA
insert collective-permutes (``fdtd/initialization.py:598-611``, ``core/jax/sharding.py:160-176``).
Here rank r owns planes ``[x0, x1)``; before the E half-step it needs ``Hy, Hz`` of plane ``x0-1``
(backward difference, ``curl.py:360-361``) and before the H half-step ``Ey, Ez`` of plane ``x1``
(forward difference, ``curl.py:273-274``): two tangential components of one plane, to/from one
neighbour, per half-step.  There is no collective in the step.

Two transports:

* ``halo="peer"`` (default on CUDA): nothing is exchanged.  Every rank maps its neighbours' own E / H
  arrays through CUDA IPC over NVLink and the half-step kernels read the boundary plane in place
  (TMA tile loads for the H half-step, 128-bit loads for the E half-step's register queue); two
  stream-ordered progress counters per rank keep producer/consumer order (``csrc/abi.cu``:
  ``peer_wait_kernel`` / ``peer_signal_kernel``).  A whole multi-step run is ONE asynchronous call
  into the C ABI per rank; ``torch.distributed`` is only used to ship the 64-byte IPC handles.
* ``halo="nccl"``: packed planes over ``torch.distributed`` send/recv (NCCL on GPUs, gloo in the CPU
  tests).  The interior planes of a half-step do not depend on the halo, so each half-step is issued
  as ``interior`` (all x-chunks but the one touching the neighbour) while the halo plane travels on a
  side stream, then the ``edge`` chunk once it has landed.
"""

from __future__ import annotations

import ctypes as C

from fdtdx_b200 import _lib
from fdtdx_b200._lib import check
from fdtdx_b200.plan import Plan


def slab_bounds(nx: int, world: int, rank: int) -> tuple[int, int]:
    """Equal x-slabs; like the reference (``sharding.py:169-173``) ``nx`` must divide evenly."""
    if nx % world != 0:
        raise ValueError(f"Nx={nx} must be divisible by the number of ranks ({world})")
    w = nx // world
    return rank * w, (rank + 1) * w


def neighbours(rank: int, world: int, periodic_x: bool) -> tuple[int | None, int | None]:
    lo = rank - 1 if rank > 0 else (world - 1 if periodic_x and world > 1 else None)
    hi = rank + 1 if rank < world - 1 else (0 if periodic_x and world > 1 else None)
    return lo, hi


def _state_x_axis(det, local, key):
    """Axis of a detector-state array (time axis included) that runs along the region's x extent, or None
    for states without one (volume reductions, the YZ slice plane)."""
    g = det._shape_dtype_single_time_step()[key][0]
    l = local._shape_dtype_single_time_step()[key][0]
    diff = [i for i, (a, b) in enumerate(zip(g, l)) if a != b]
    return None if not diff else diff[0] + 1


def merge_detector_states(objects, config, bounds, states_per_rank):
    """Whole-region detector states from the per-rank parts of an x-sharded run (``bounds``: the ranks'
    ``(x0, x1)``; ``states_per_rank``: their ``detector_states`` dicts as NumPy arrays).  Region-shaped
    states are concatenated along x; sums over the region (Poynting / energy ``reduce_volume``, weighted
    means normalised by the global weight sum) add up; the YZ mean of an energy slice set is the
    plane-count-weighted mean of the ranks' means (one reduction after the run)."""
    import numpy as np

    from fdtdx_b200.detectors import EnergyDetector
    from fdtdx_b200.plan import clip_detector

    out = {}
    for det in objects.detectors:
        dlo, dhi = det.grid_slice_tuple[0]
        parts = [(r, clip_detector(det, x0, x1, config)) for r, (x0, x1) in enumerate(bounds) if dhi > x0 and dlo < x1]
        if len(parts) == 1:
            out[det.name] = {k: np.asarray(v) for k, v in states_per_rank[parts[0][0]][det.name].items()}
            continue
        merged = {}
        for key in states_per_rank[parts[0][0]][det.name]:
            ax = _state_x_axis(det, parts[0][1], key)
            vals = [np.asarray(states_per_rank[r][det.name][key]) for r, _ in parts]
            if ax is not None:
                merged[key] = np.concatenate(vals, axis=ax)
            elif isinstance(det, EnergyDetector) and det.as_slices and det.use_mean and key == "YZ Plane":
                ex = float(dhi - dlo)
                merged[key] = sum(v * np.float32((loc.grid_slice_tuple[0][1] - loc.grid_slice_tuple[0][0]) / ex) for v, (_, loc) in zip(vals, parts)).astype(vals[0].dtype)
            else:
                merged[key] = sum(vals[1:], vals[0]).astype(vals[0].dtype)
        out[det.name] = merged
    return out


def shard_arrays(arrays, objects, x_range, device=None, config=None):
    """This rank's x-slab of a whole-grid (NumPy) ``ArrayContainer`` - what the reference's
    ``NamedSharding`` on the x axis gives every device (``fdtd/initialization.py:598-611``,
    ``interfaces/state.py:72-78``): fields, materials, ADE arrays and y / z CPML psi are sliced on x; an x
    CPML slab and an x interface plane of the recorder stay with the rank that owns their planes; a
    detector whose region straddles a slab edge keeps, on every rank, the state of its part of the region
    (needs ``config``; ``merge_detector_states`` puts them back together).  ``device``: move to that torch device."""
    import numpy as np

    from fdtdx_b200.container import ArrayContainer, FieldState, RecordingState, _TorchLeaf

    x0, x1 = x_range

    def cut(a, axis):
        if a is None or not hasattr(a, "shape") or len(a.shape) == 0:
            return a
        idx = [slice(None)] * len(a.shape)
        idx[axis] = slice(x0, x1)
        if isinstance(a, _TorchLeaf):
            return _TorchLeaf(a.t[tuple(idx)].contiguous())
        return np.ascontiguousarray(a[tuple(idx)])

    f = arrays.fields
    psi_E, psi_H = {}, {}
    for pml in objects.pml_objects:
        lo, hi = pml.grid_slice_tuple[0]
        a, b = max(lo, x0), min(hi, x1)
        if b <= a:
            continue
        sl = (slice(a - lo, b - lo),)
        psi_E[pml.name] = tuple(np.ascontiguousarray(q[sl]) for q in f.psi_E[pml.name])
        psi_H[pml.name] = tuple(np.ascontiguousarray(q[sl]) for q in f.psi_H[pml.name])
    det = {}
    for d in objects.detectors:
        lo, hi = d.grid_slice_tuple[0]
        if hi <= x0 or lo >= x1:
            continue
        if lo < x0 or hi > x1:
            if config is None:
                raise NotImplementedError(f"detector {d.name!r} straddles the slab edge at x = {x0 if lo < x0 else x1}: pass config to shard_arrays")
            from fdtdx_b200.plan import clip_detector

            local = clip_detector(d, x0, x1, config)
            a, b = max(lo, x0) - lo, min(hi, x1) - lo
            st = {}
            for key, v in arrays.detector_states[d.name].items():
                ax = _state_x_axis(d, local, key)
                if ax is None:
                    st[key] = np.zeros_like(v) if (a > 0) else np.array(v)  # sums: the whole-region value counts once
                else:
                    idx = [slice(None)] * v.ndim
                    idx[ax] = slice(a, b)
                    st[key] = np.ascontiguousarray(v[tuple(idx)])
            det[d.name] = st
            continue
        det[d.name] = arrays.detector_states[d.name]
    rec = None
    if arrays.recording_state is not None:
        data = {}
        for pml in objects.pml_objects:
            for fs in ("E", "H"):
                buf = arrays.recording_state.data[f"{pml.name}_{fs}"]
                if pml.axis == 0:
                    plane = pml.interface_slice()[0].start
                    if x0 <= plane < x1:
                        data[f"{pml.name}_{fs}"] = buf
                else:
                    data[f"{pml.name}_{fs}"] = cut(buf, 2)  # (slots, 3, Nx, ...)
        rec = RecordingState(data=data, state={})
    out = ArrayContainer(
        fields=FieldState(E=cut(f.E, 1), H=cut(f.H, 1), psi_E=psi_E, psi_H=psi_H, dispersive_P_curr=cut(f.dispersive_P_curr, 2), dispersive_P_prev=cut(f.dispersive_P_prev, 2)),
        inv_permittivities=cut(arrays.inv_permittivities, 1), inv_permeabilities=cut(arrays.inv_permeabilities, 1), detector_states=det, recording_state=rec,
        electric_conductivity=cut(arrays.electric_conductivity, 1), magnetic_conductivity=cut(arrays.magnetic_conductivity, 1),
        dispersive_c1=cut(arrays.dispersive_c1, 2), dispersive_c2=cut(arrays.dispersive_c2, 2), dispersive_c3=cut(arrays.dispersive_c3, 2), dispersive_c4=cut(arrays.dispersive_c4, 2),
    )
    return out if device is None else out.to_torch(device)


class HaloExchange:
    """Point-to-point exchange of one packed plane with the two x-neighbours over
    ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests)."""

    def __init__(self, rank: int, world: int, periodic_x: bool, group=None):
        self.rank, self.world = rank, world
        self.lo, self.hi = neighbours(rank, world, periodic_x)
        self.group = group

    def exchange(self, send_to_hi, recv_from_lo, send_to_lo, recv_from_hi):
        """Any argument may be None.  Returns after the receives have completed (on the current
        CUDA stream for NCCL)."""
        import torch.distributed as dist

        ops = []
        if send_to_hi is not None and self.hi is not None:
            ops.append(dist.P2POp(dist.isend, send_to_hi, self.hi, self.group))
        if recv_from_lo is not None and self.lo is not None:
            ops.append(dist.P2POp(dist.irecv, recv_from_lo, self.lo, self.group))
        if send_to_lo is not None and self.lo is not None:
            ops.append(dist.P2POp(dist.isend, send_to_lo, self.lo, self.group))
        if recv_from_hi is not None and self.hi is not None:
            ops.append(dist.P2POp(dist.irecv, recv_from_hi, self.hi, self.group))
        if not ops:
            return
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class SlabRunner:
    """Drives one rank's slab through the C ABI with halo exchange between the half-steps."""

    def __init__(self, objects, config, arrays, x_range, rank: int, world: int, group=None, overlap: bool = True, halo: str | None = None):
        import os

        import torch

        self.objects, self.config, self.arrays = objects, config, arrays
        self.rank, self.world = rank, world
        periodic_x = any(b.uses_wrap_padding and b.axis == 0 for b in objects.boundary_objects)
        self.hx = HaloExchange(rank, world, periodic_x, group)
        has_lo, has_hi = self.hx.lo is not None, self.hx.hi is not None
        self.plan = Plan(objects, config, arrays, x_range=x_range, halo=(has_lo, has_hi))
        self.plan.bind(arrays)
        E, H = arrays.fields.E, arrays.fields.H
        ny, nz = E.shape[2], E.shape[3]
        # Exact-interpolation detectors whose region continues across a slab edge read plane x-1 of the lower
        # neighbour (co-location stencil, curl.py:86-224): on the steps they are on, that plane of H (before
        # the H update), E and H (after) travels to the upper neighbour
        x0, x1 = x_range
        nxg = objects.volume.grid_shape[0]
        crosses = lambda d, X: d.exact_interpolation and d.grid_slice_tuple[0][0] < X < d.grid_slice_tuple[0][1]
        wraps = lambda d: periodic_x and d.exact_interpolation and d.grid_slice_tuple[0][0] == 0  # x-1 of plane 0 wraps to the last rank
        self.det_from_lo = [d for d in objects.detectors if has_lo and (crosses(d, x0) or (x0 == 0 and wraps(d)))]
        self.det_to_hi = [d for d in objects.detectors if has_hi and (crosses(d, x1) or (x1 == nxg and wraps(d)))]
        if self.det_from_lo or self.det_to_hi:
            mk3 = lambda: torch.zeros((3, ny, nz), dtype=torch.float32, device=E.device)
            self.xlo_E, self.xlo_H, self.xlo_Hp, self.send3a, self.send3b = mk3(), mk3(), mk3(), mk3(), mk3()
            if self.det_from_lo:
                self.plan._bind(_lib.SLOT_DET_XLO_E, 0, self.xlo_E)
                self.plan._bind(_lib.SLOT_DET_XLO_H, 0, self.xlo_H)
                self.plan._bind(_lib.SLOT_DET_XLO_HPREV, 0, self.xlo_Hp)
        mk = lambda: torch.zeros((2, ny, nz), dtype=torch.float32, device=E.device)
        self.haloH, self.haloE, self.sendH, self.sendE = mk(), mk(), mk(), mk()
        if has_lo:
            self.plan._bind(_lib.SLOT_HALO_H_LO, 0, self.haloH)
        if has_hi:
            self.plan._bind(_lib.SLOT_HALO_E_HI, 0, self.haloE)
        self.overlap = overlap and world > 1
        self.nx = E.shape[1]
        halo = halo or os.environ.get("FDTDX_B200_HALO", "peer")
        self.peer = bool(halo == "peer" and E.is_cuda and world > 1 and (has_lo or has_hi))
        if self.peer:
            self._attach_peers(group)
        self.side = torch.cuda.Stream(device=E.device) if (self.overlap and E.is_cuda and not self.peer) else None  # after the attach: it may have fallen back

    def _attach_peers(self, group):
        """Ship this rank's IPC handles (E, H, progress flags) to everybody, attach the two neighbours'.
        If any rank cannot export or map (no CUDA IPC in the container, expandable-segment allocator, no
        peer access), every rank falls back to the exchanged-plane transport together."""
        import warnings

        import torch.distributed as dist

        lib, h = self.plan.lib, self.plan.h
        import os

        mine = {"nx": int(self.nx), "ok": True}
        try:
            if os.environ.get("FDTDX_B200_PEER_FAIL") == str(self.rank):  # test hook for the fallback path
                raise RuntimeError("simulated CUDA IPC failure (FDTDX_B200_PEER_FAIL)")
            for what, name in ((0, "E"), (1, "H"), (2, "flags")):
                buf = C.create_string_buffer(64)
                off = C.c_longlong(0)
                check(lib.fdtdx_b200_peer_export(h, what, buf, C.byref(off)))
                mine[name] = (bytes(buf.raw), int(off.value))
        except RuntimeError as e:
            mine = {"nx": int(self.nx), "ok": False, "why": str(e)}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        ok = all(q["ok"] for q in everyone)
        why = next((q.get("why") for q in everyone if not q["ok"]), None)
        if ok:
            try:
                if self.hx.lo is not None:
                    q = everyone[self.hx.lo]
                    check(lib.fdtdx_b200_peer_attach(h, 0, q["H"][0], q["H"][1], q["flags"][0], q["flags"][1], q["nx"]))
                if self.hx.hi is not None:
                    q = everyone[self.hx.hi]
                    check(lib.fdtdx_b200_peer_attach(h, 1, q["E"][0], q["E"][1], q["flags"][0], q["flags"][1], q["nx"]))
            except RuntimeError as e:
                ok, why = False, str(e)
        verdict = [None] * self.world
        dist.all_gather_object(verdict, bool(ok), group=group)
        if not all(verdict):
            check(lib.fdtdx_b200_peer_detach(h))
            self.peer = False
            if self.rank == 0:
                warnings.warn(f"fdtdx_b200: peer-memory halo unavailable ({why}); using the NCCL plane exchange")
            return
        dist.barrier(group=group)  # nobody starts stepping before every mapping exists

    def _range(self, t, which, x_begin, x_end, simulate=True):
        import torch

        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self.plan.lib.fdtdx_b200_run_half_range(self.plan.h, int(t), int(which), int(x_begin), int(x_end), int(simulate), st))

    def _det_halo_on(self, t: int, record_detectors: bool, inverse: bool = False):
        on = lambda ds: record_detectors and any(bool(d._is_on_at_time_step_arr[t]) and bool(d.inverse) == inverse for d in ds)
        return on(self.det_from_lo), on(self.det_to_hi)

    def _det_halo(self, t: int, record_detectors: bool, before_H: bool):
        """Plane x-1 for the straddling detectors that are on at step t: H before the H update, or E and H
        after it, from every rank to its upper neighbour (the same on/off tables on both sides of an edge)."""
        recv, send = self._det_halo_on(t, record_detectors)
        if not (recv or send):
            return
        E, H = self.arrays.fields.E, self.arrays.fields.H
        if before_H:
            if send:
                self.send3a.copy_(H[:, -1])
            self.hx.exchange(self.send3a if send else None, self.xlo_Hp if recv else None, None, None)
        else:
            if send:
                self.send3a.copy_(E[:, -1])
                self.send3b.copy_(H[:, -1])
            self.hx.exchange(self.send3a if send else None, self.xlo_E if recv else None, None, None)
            self.hx.exchange(self.send3b if send else None, self.xlo_H if recv else None, None, None)

    def step(self, t: int, record_detectors: bool = False, record_boundaries: bool = False):
        import torch

        E, H = self.arrays.fields.E, self.arrays.fields.H
        if self.peer:
            self._det_halo(t, record_detectors, True)
            for phase in range(3):
                if phase == 2:
                    self._det_halo(t, record_detectors, False)
                self.plan.run_forward_phase(t, phase, record_detectors, record_boundaries, True)
            return
        self._det_halo(t, record_detectors, True)
        xc = self.plan.xchunk_hint()
        if self.side is None:
            self.sendH.copy_(H[1:3, -1])
            self.hx.exchange(self.sendH, self.haloH, None, None)
            self.plan.run_forward_phase(t, 0, record_detectors, False, True)
            self.sendE.copy_(E[1:3, 0])
            self.hx.exchange(None, None, self.sendE, self.haloE)
            self.plan.run_forward_phase(t, 1, record_detectors, False, True)
        else:
            main = torch.cuda.current_stream()
            # --- E half-step: needs Hy,Hz of x0-1 only in the first chunk
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                self.sendH.copy_(H[1:3, -1])
                self.hx.exchange(self.sendH, self.haloH, None, None)
            if record_detectors:
                self.plan.run_forward_phase(t, 3, True, False, True)  # detector H_prev gather only
            self._range(t, 0, xc, self.nx)
            main.wait_stream(self.side)
            self._range(t, 0, 0, xc)
            # --- H half-step: needs Ey,Ez of x1 only in the last chunk
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                self.sendE.copy_(E[1:3, 0])
                self.hx.exchange(None, None, self.sendE, self.haloE)
            last = max(self.nx - xc, 0)
            self._range(t, 1, 0, last)
            main.wait_stream(self.side)
            self._range(t, 1, last, self.nx)
        self._det_halo(t, record_detectors, False)
        self.plan.run_forward_phase(t, 2, record_detectors, False, True)

    def run(self, t0: int, n: int, record_detectors: bool = False, record_boundaries: bool = False):
        if self.peer:
            # one asynchronous submission per stretch of steps on which no straddling detector is on; the
            # steps in between are issued phase by phase around the detector-plane exchange
            t, end = int(t0), int(t0) + int(n)
            while t < end:
                m = 0
                while t + m < end and not any(self._det_halo_on(t + m, record_detectors)):
                    m += 1
                if m:
                    self.plan.run_forward(t, m, record_detectors, record_boundaries, True)
                    t += m
                else:
                    self.step(t, record_detectors, record_boundaries)
                    t += 1
            return
        if record_boundaries:
            raise NotImplementedError("interface recording on slabs runs on the peer-memory halo")
        for t in range(t0, t0 + n):
            self.step(t, record_detectors)

    def gather_detector_states(self, group=None):
        """Whole-region detector states on every rank (``merge_detector_states``) - the one reduction after
        the run that SURVEY section 8e allows."""
        import torch.distributed as dist

        mine = {d: {k: v.detach().cpu().numpy() for k, v in st.items()} for d, st in self.arrays.detector_states.items()}
        if self.world == 1:
            return mine
        x0 = self.plan.x0
        everyone = [None] * self.world
        dist.all_gather_object(everyone, ((x0, self.plan.x1), mine), group=group)
        return merge_detector_states(self.objects, self.config, [b for b, _ in everyone], [s for _, s in everyone])

    def run_reverse(self, t_from: int, n: int, record_detectors: bool = False, reset_fields: bool = True):
        """``full_backward`` on this rank's slab (``backward.py:18-135``): interface replay, reverse H and
        E half-steps reading the neighbours' planes in place, PML field reset, inverse detectors.  One
        asynchronous submission; needs the peer-memory halo."""
        if not self.peer:
            raise NotImplementedError("the time-reversed pass on slabs runs on the peer-memory halo")
        if record_detectors and any(d.inverse for d in self.det_from_lo + self.det_to_hi):
            raise NotImplementedError("inverse detectors straddling a slab edge in the time-reversed pass")
        self.plan.run_reverse(t_from, n, record_detectors, reset_fields)
