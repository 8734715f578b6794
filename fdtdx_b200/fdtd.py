"""Loop drivers with the reference's signatures, running on the sm_100a kernels.

Mirrors ``fdtdx/fdtd/wrapper.py:14-63`` (``run_fdtd``), ``fdtdx/fdtd/fdtd.py:39-584``
(``reversible_fdtd``, ``checkpointed_fdtd``, ``custom_fdtd_forward``),
``fdtdx/fdtd/forward.py:83-156`` (``forward``) and ``fdtdx/fdtd/backward.py:18-135``
(``full_backward``, ``backward``).  Arrays are CUDA ``torch.Tensor`` leaves of an
``ArrayContainer``; like the reference under ``donate_argnames`` the field/state buffers are
updated in place and the returned container aliases them.  There is no CPU path: without the
CUDA extension every driver raises.
"""

from __future__ import annotations

from typing import Callable

from fdtdx_b200.config import SimulationConfig
from fdtdx_b200.container import ArrayContainer, ObjectContainer, SimulationState
from fdtdx_b200.plan import Plan


def _require_cuda(arrays: ArrayContainer):
    import torch

    E = arrays.fields.E
    if not isinstance(E, torch.Tensor) or not E.is_cuda:
        raise RuntimeError(
            "fdtdx_b200 runs on CUDA tensors only (no CPU fallback): move the container with arrays.to_torch('cuda')"
        )


def get_plan(arrays: ArrayContainer, objects: ObjectContainer, config: SimulationConfig) -> Plan:
    """Plan cache on the ObjectContainer (plans hold only constant tables + scratch)."""
    _require_cuda(arrays)
    cache = objects.__dict__.setdefault("_plan_cache", {})
    mu = arrays.inv_permeabilities
    # the cache entry keeps the config alive (Plan.config), so its id cannot be recycled while the
    # entry exists; entries whose config is a different object are never returned
    from fdtdx_b200.bloch import ComplexPlan, is_complex_run

    cplx = is_complex_run(arrays)
    key = (
        id(config),
        cplx,
        arrays.fields.E.device.index,
        tuple(arrays.inv_permittivities.shape),
        tuple(mu.shape) if hasattr(mu, "shape") else float(mu),
        None if arrays.electric_conductivity is None else tuple(arrays.electric_conductivity.shape),
        None if arrays.magnetic_conductivity is None else tuple(arrays.magnetic_conductivity.shape),
        None if arrays.dispersive_c1 is None else tuple(arrays.dispersive_c1.shape),
        arrays.dispersive_c4 is not None,
    )
    plan = cache.get(key)
    if plan is not None and plan.config is not config:
        plan = None
    if plan is None:
        # complex fields (Bloch k != 0, initialization.py:581-596): two lock-stepped real systems
        plan = ComplexPlan(objects, config, arrays) if cplx else Plan(objects, config, arrays)
        cache[key] = plan
    plan.bind(arrays)
    return plan


def _progress(show_progress: bool, progress_callback, t: int, total: int):
    if progress_callback is not None:
        progress_callback(t, total)


def _run_forward_loop(arrays, objects, config, start: int, end: int, record_detectors: bool, record_boundaries: bool, progress_callback=None):
    plan = get_plan(arrays, objects, config)
    n = int(end) - int(start)
    if n > 0:
        if progress_callback is None:
            plan.run_forward(int(start), n, record_detectors, record_boundaries, True)
        else:
            chunk = max(1, n // 100)
            t = int(start)
            while t < end:
                m = min(chunk, int(end) - t)
                plan.run_forward(t, m, record_detectors, record_boundaries, True)
                t += m
                progress_callback(t, int(end))
    return plan.finish(arrays)


def checkpointed_fdtd(
    arrays: ArrayContainer,
    objects: ObjectContainer,
    config: SimulationConfig,
    key=None,
    stopping_condition=None,
    show_progress: bool = True,
    progress_callback: Callable[[int, int], None] | None = None,
) -> SimulationState:
    """``fdtd.py:421-496``: reset, then ``forward`` until ``time_steps_total``.

    With a ``stopping_condition`` (``fdtd/stop_conditions.py``) the loop is the reference's
    ``while cond(state): state = forward(state)`` bounded by ``time_steps_total`` (``fdtd.py:482-493``):
    steps are issued in bulk while the condition cannot fire (``earliest_stop``) and one at a time,
    with one device reduction + 4-byte read-back each, afterwards."""
    T = config.time_steps_total
    gc = config.gradient_config
    if stopping_condition is None and gc is not None and gc.method == "checkpointed":
        import torch

        inv_eps, inv_mu = arrays.inv_permittivities, arrays.inv_permeabilities
        coef_grad = any(torch.is_tensor(c) and c.requires_grad for c in (arrays.dispersive_c1, arrays.dispersive_c2, arrays.dispersive_c3, arrays.dispersive_c4))
        needs_grad = torch.is_tensor(inv_eps) and (inv_eps.requires_grad or (isinstance(inv_mu, torch.Tensor) and inv_mu.requires_grad) or coef_grad)
        if needs_grad and torch.is_grad_enabled():
            return _checkpointed_with_grad(arrays, objects, config, progress_callback)
    arrays = arrays.reset()
    if stopping_condition is None:
        arrays = _run_forward_loop(arrays, objects, config, 0, T, True, config.invertible_optimization, progress_callback)
        return T, arrays
    cond = stopping_condition.setup((0, arrays), config, objects)
    t = min(max(int(cond.earliest_stop(config)), 0), T)
    if t > 0:
        if not cond((0, arrays), config, objects):  # a condition that refuses the initial state
            return 0, arrays
        arrays = _run_forward_loop(arrays, objects, config, 0, t, True, config.invertible_optimization, progress_callback)
    while t < T and cond((t, arrays), config, objects):
        arrays = _run_forward_loop(arrays, objects, config, t, t + 1, True, config.invertible_optimization, None)
        t += 1
        if progress_callback is not None:
            progress_callback(t, T)
    return t, arrays


def custom_fdtd_forward(
    arrays: ArrayContainer,
    objects: ObjectContainer,
    config: SimulationConfig,
    key=None,
    reset_container: bool = True,
    record_detectors: bool = True,
    start_time: int = 0,
    end_time: int = 0,
    show_progress: bool = True,
    progress_callback: Callable[[int, int], None] | None = None,
) -> SimulationState:
    """``fdtd.py:499-584``."""
    if reset_container:
        arrays = arrays.reset()
    end = max(int(end_time), int(start_time))
    arrays = _run_forward_loop(arrays, objects, config, int(start_time), end, record_detectors, False, progress_callback)
    return end, arrays


def _checkpointed_with_grad(arrays, objects, config, progress_callback):
    """``GradientConfig(method="checkpointed")`` (``fdtd.py:482-493``, ``kind="checkpointed"``): the
    forward run keeps ``num_checkpoints`` full states; the backward pass recomputes each segment from
    its checkpoint, keeps that segment's per-step states, and applies the fused adjoint kernels at the
    *stored* states (``fdtdx_b200_run_adjoint_exact``) - no time reversal, so the gradient is that of
    the forward run itself, CPML included.  Differentiable w.r.t. the materials, like the reference's
    ``reversible_fdtd``; dispersive media are not supported by the adjoint kernels yet."""
    import torch

    _require_cuda(arrays)
    inv_eps, inv_mu = arrays.inv_permittivities, arrays.inv_permeabilities
    holder = {"arrays": arrays, "objects": objects, "config": config, "cb": progress_callback}
    coeffs = [arrays.dispersive_c1, arrays.dispersive_c2, arrays.dispersive_c3, arrays.dispersive_c4]
    outs = _CheckpointedFunction.get().apply(inv_eps, inv_mu if isinstance(inv_mu, torch.Tensor) else None, holder, *coeffs)
    out = holder["out"]
    out = out.aset("fields->E", outs[0]).aset("fields->H", outs[1])
    det = {d: dict(st) for d, st in out.detector_states.items()}
    for (d, k), v in zip(holder["names"], outs[2:]):
        det[d][k] = v
    out = out.aset("detector_states", det)
    out = out.aset("inv_permittivities", inv_eps)
    if isinstance(inv_mu, torch.Tensor):
        out = out.aset("inv_permeabilities", inv_mu)
    for name, c in zip(("dispersive_c1", "dispersive_c2", "dispersive_c3", "dispersive_c4"), coeffs):
        if c is not None:
            out = out.aset(name, c)
    return config.time_steps_total, out


def _detector_cotangents(names, gdet, arrays):
    """Detector-state cotangents as plain contiguous buffers the kernels can read by pointer.  Autograd
    may hand over a complex gradient with the lazy conj / neg bit set (e.g. a loss built from
    ``phasor.conj()``); ``contiguous()`` alone keeps the bit, so it is materialised here."""
    cot_det = {}
    for (d, k), g in zip(names, gdet):
        if g is None:
            continue
        st = arrays.detector_states[d][k]
        g = g.detach().resolve_conj().resolve_neg().contiguous()
        if g.dtype != st.dtype or tuple(g.shape) != tuple(st.shape):
            raise ValueError(f"cotangent of detector state {d}/{k}: {g.dtype}{tuple(g.shape)} != state {st.dtype}{tuple(st.shape)}")
        cot_det.setdefault(d, {})[k] = g
    return cot_det


def _clone_state(a):
    f = a.fields
    P = None if f.dispersive_P_curr is None else (f.dispersive_P_curr.clone(), f.dispersive_P_prev.clone())
    return (f.E.clone(), f.H.clone(), {k: (x.clone(), y.clone()) for k, (x, y) in f.psi_E.items()}, {k: (x.clone(), y.clone()) for k, (x, y) in f.psi_H.items()}, P)


def _load_state(a, st):
    E, H, pE, pH, P = st
    a.fields.E.copy_(E)
    a.fields.H.copy_(H)
    for k in pE:
        for w in range(2):
            a.fields.psi_E[k][w].copy_(pE[k][w])
            a.fields.psi_H[k][w].copy_(pH[k][w])
    if P is not None:
        a.fields.dispersive_P_curr.copy_(P[0])
        a.fields.dispersive_P_prev.copy_(P[1])


class _CheckpointedFunction:
    _cls = None

    @classmethod
    def get(cls):
        if cls._cls is not None:
            return cls._cls
        import torch

        class CheckpointedFDTD(torch.autograd.Function):
            @staticmethod
            def forward(ctx, inv_eps, inv_mu, holder, c1=None, c2=None, c3=None, c4=None):
                arrays, objects, config = holder["arrays"], holder["objects"], holder["config"]
                T = config.time_steps_total
                nck = max(1, int(config.gradient_config.num_checkpoints))
                seg_len = max(1, -(-T // nck))
                holder["coeff_needs_grad"] = [c is not None and c.requires_grad for c in (c1, c2, c3, c4)]
                with torch.no_grad():
                    arrays = arrays.aset("inv_permittivities", inv_eps.detach())
                    if inv_mu is not None:
                        arrays = arrays.aset("inv_permeabilities", inv_mu.detach())
                    for name, c in zip(("dispersive_c1", "dispersive_c2", "dispersive_c3", "dispersive_c4"), (c1, c2, c3, c4)):
                        if c is not None:
                            arrays = arrays.aset(name, c.detach())
                    arrays = arrays.reset()
                    ckpts, t = [], 0
                    while t < T:
                        ckpts.append((t, _clone_state(arrays)))
                        m = min(seg_len, T - t)
                        arrays = _run_forward_loop(arrays, objects, config, t, t + m, True, False, holder["cb"])
                        t += m
                holder["ckpts"], holder["out"] = ckpts, arrays
                ctx.holder = holder
                ctx.set_materialize_grads(False)
                names = [(d, k) for d, st in arrays.detector_states.items() for k in st]
                holder["names"] = names
                return (arrays.fields.E, arrays.fields.H, *[arrays.detector_states[d][k] for d, k in names])

            @staticmethod
            def backward(ctx, gE, gH, *gdet):
                h = ctx.holder
                arrays, objects, config = h["out"], h["objects"], h["config"]
                T = config.time_steps_total
                f = arrays.fields
                work = arrays.aset("fields->E", f.E.detach().clone()).aset("fields->H", f.H.detach().clone())
                work = work.aset("fields->psi_E", {k: (a.clone(), b.clone()) for k, (a, b) in f.psi_E.items()})
                work = work.aset("fields->psi_H", {k: (a.clone(), b.clone()) for k, (a, b) in f.psi_H.items()})
                cot_P = cot_Q = g_coef = None
                if f.dispersive_P_curr is not None:
                    # the final P / P_prev are not outputs of the autograd node: their cotangents start at zero
                    work = work.aset("fields->dispersive_P_curr", f.dispersive_P_curr.clone()).aset("fields->dispersive_P_prev", f.dispersive_P_prev.clone())
                    cot_P, cot_Q = torch.zeros_like(f.dispersive_P_curr), torch.zeros_like(f.dispersive_P_prev)
                    cs = (work.dispersive_c1, work.dispersive_c2, work.dispersive_c3, work.dispersive_c4)
                    g_coef = [torch.zeros_like(c) if (c is not None and need) else None for c, need in zip(cs, h["coeff_needs_grad"])]
                cot_E = torch.zeros_like(f.E) if gE is None else gE.detach().clone().contiguous()
                cot_H = torch.zeros_like(f.H) if gH is None else gH.detach().clone().contiguous()
                cot_det = _detector_cotangents(h["names"], gdet, arrays)
                g_eps = torch.zeros_like(work.inv_permittivities)
                mu = work.inv_permeabilities
                g_mu = torch.zeros_like(mu) if isinstance(mu, torch.Tensor) else None
                plan = get_plan(work, objects, config)
                first = True
                ends = [c[0] for c in h["ckpts"][1:]] + [T]
                for (t0, st0), t1 in zip(reversed(h["ckpts"]), reversed(ends)):
                    # recompute the segment from its checkpoint, keeping every step's input state
                    _load_state(work, st0)
                    states = []
                    for t in range(t0, t1):
                        states.append(_clone_state(work))
                        plan.bind(work)
                        plan.run_forward(t, 1, False, False, True)
                        work = plan.finish(work)  # ADE: P_curr / P_prev swap roles every step
                    for t in range(t1 - 1, t0 - 1, -1):
                        _load_state(work, states[t - t0])
                        plan.run_adjoint(work, t + 1, 1, cot_E, cot_H, cot_det, g_eps, g_mu, keep_cot_psi=not first, exact=True,
                                         cot_P=cot_P, cot_P_prev=cot_Q, grad_coeffs=g_coef)
                        first = False
                    del states
                gc = [None] * 4 if g_coef is None else g_coef
                return (g_eps, g_mu, None, *gc)

        cls._cls = CheckpointedFDTD
        return CheckpointedFDTD


class _ReversibleFunction:
    """Built lazily so that importing this module does not import torch."""

    _cls = None

    @classmethod
    def get(cls):
        if cls._cls is not None:
            return cls._cls
        import torch

        class ReversibleFDTD(torch.autograd.Function):
            """custom_vjp of ``reversible_fdtd`` (``fdtd/fdtd.py:176-379``): differentiable w.r.t.
            ``inv_permittivities`` / ``inv_permeabilities`` only (``:324-333``)."""

            @staticmethod
            def forward(ctx, inv_eps, inv_mu, holder):
                arrays, objects, config, progress_callback = holder["arrays"], holder["objects"], holder["config"], holder["cb"]
                T = config.time_steps_total
                with torch.no_grad():
                    arrays = arrays.aset("inv_permittivities", inv_eps.detach())
                    if inv_mu is not None:
                        arrays = arrays.aset("inv_permeabilities", inv_mu.detach())
                    arrays = arrays.reset()
                    # sliced reversible pass (fdtd.py:106-166): full FieldState checkpoints at the interior
                    # slice boundaries s_1..s_{k-1}
                    bounds = holder["bounds"]
                    ckpts = []
                    for seg in range(len(bounds) - 1):
                        arrays = _run_forward_loop(arrays, objects, config, bounds[seg], bounds[seg + 1], True, config.invertible_optimization, progress_callback)
                        if seg < len(bounds) - 2:
                            f = arrays.fields
                            ckpts.append(
                                (f.E.clone(), f.H.clone(), {k: (a.clone(), b.clone()) for k, (a, b) in f.psi_E.items()}, {k: (a.clone(), b.clone()) for k, (a, b) in f.psi_H.items()})
                            )
                    holder["ckpts"] = ckpts
                holder["out"] = arrays
                ctx.holder = holder
                ctx.set_materialize_grads(False)
                names = [(d, k) for d, st in arrays.detector_states.items() for k in st]
                holder["names"] = names
                outs = (arrays.fields.E, arrays.fields.H, *[arrays.detector_states[d][k] for d, k in names])
                return outs

            @staticmethod
            def backward(ctx, gE, gH, *gdet):
                h = ctx.holder
                arrays, objects, config = h["out"], h["objects"], h["config"]
                if not config.invertible_optimization:
                    raise Exception("Need recorder to record boundaries")
                T = config.time_steps_total
                # reconstruct on copies so that the user's final fields stay intact (JAX is functional)
                work = arrays.aset("fields->E", arrays.fields.E.detach().clone()).aset("fields->H", arrays.fields.H.detach().clone())
                cot_E = torch.zeros_like(work.fields.E) if gE is None else gE.detach().clone().contiguous()
                cot_H = torch.zeros_like(work.fields.H) if gH is None else gH.detach().clone().contiguous()
                cot_det = _detector_cotangents(h["names"], gdet, arrays)
                g_eps = torch.zeros_like(work.inv_permittivities)
                mu = work.inv_permeabilities
                g_mu = torch.zeros_like(mu) if isinstance(mu, torch.Tensor) else None
                plan = get_plan(work, objects, config)
                bounds, ckpts = h["bounds"], h["ckpts"]
                if ckpts:
                    # psi is restored together with the fields at each boundary, so work on copies of it too
                    work = work.aset("fields->psi_E", {k: (a.clone(), b.clone()) for k, (a, b) in work.fields.psi_E.items()})
                    work = work.aset("fields->psi_H", {k: (a.clone(), b.clone()) for k, (a, b) in work.fields.psi_H.items()})
                for seg in range(len(bounds) - 2, -1, -1):
                    plan.run_adjoint(work, bounds[seg + 1], bounds[seg + 1] - bounds[seg], cot_E, cot_H, cot_det, g_eps, g_mu, keep_cot_psi=(seg != len(bounds) - 2))
                    if seg > 0:
                        # reset the reconstructed primal state to the exact checkpoint at s_i (fdtd.py:300-313)
                        cE, cH, cpE, cpH = ckpts[seg - 1]
                        work.fields.E.copy_(cE)
                        work.fields.H.copy_(cH)
                        for k in cpE:
                            for w in range(2):
                                work.fields.psi_E[k][w].copy_(cpE[k][w])
                                work.fields.psi_H[k][w].copy_(cpH[k][w])
                return g_eps, g_mu, None

        cls._cls = ReversibleFDTD
        return ReversibleFDTD


def reversible_fdtd(
    arrays: ArrayContainer,
    objects: ObjectContainer,
    config: SimulationConfig,
    key=None,
    show_progress: bool = True,
    progress_callback: Callable[[int, int], None] | None = None,
) -> SimulationState:
    """``fdtd.py:39-418``: forward pass with PML-interface recording; when ``inv_permittivities`` (or
    ``inv_permeabilities``) requires grad, the result is wired into torch autograd and the backward
    pass runs the time-reversed reconstruction + fused adjoint kernels (``fdtdx_b200_run_adjoint``),
    returning gradients for the materials only, like the reference's ``custom_vjp``."""
    import torch

    if arrays.dispersive_c1 is not None or arrays.fields.dispersive_P_curr is not None:
        raise NotImplementedError(
            "Dispersive time-reversible gradient computation under active development. "
            "Use GradientConfig(method='checkpointed') instead."
        )
    num_ckpt = 0 if config.gradient_config is None else config.gradient_config.num_checkpoints_reversible
    if num_ckpt > 0 and num_ckpt + 1 > config.time_steps_total:
        raise Exception(
            "num_checkpoints_reversible must be <= time_steps_total - 1 "
            f"(got num_checkpoints_reversible={num_ckpt}, time_steps_total={config.time_steps_total})"
        )
    _require_cuda(arrays)
    T = config.time_steps_total
    inv_eps, inv_mu = arrays.inv_permittivities, arrays.inv_permeabilities
    needs_grad = inv_eps.requires_grad or (isinstance(inv_mu, torch.Tensor) and inv_mu.requires_grad)
    if not needs_grad or not torch.is_grad_enabled():
        arrays = arrays.reset()
        arrays = _run_forward_loop(arrays, objects, config, 0, T, True, config.invertible_optimization, progress_callback)
        return T, arrays
    bounds = [round(i * T / (num_ckpt + 1)) for i in range(num_ckpt + 2)]  # _reversible_slice_boundaries, fdtd.py:20-36
    holder = {"arrays": arrays, "objects": objects, "config": config, "cb": progress_callback, "bounds": bounds}
    outs = _ReversibleFunction.get().apply(inv_eps, inv_mu if isinstance(inv_mu, torch.Tensor) else None, holder)
    out = holder["out"]
    out = out.aset("fields->E", outs[0]).aset("fields->H", outs[1])
    det = {d: dict(st) for d, st in out.detector_states.items()}
    for (d, k), v in zip(holder["names"], outs[2:]):
        det[d][k] = v
    out = out.aset("detector_states", det).aset("inv_permittivities", inv_eps).aset("inv_permeabilities", inv_mu)
    return T, out


def run_fdtd(
    arrays: ArrayContainer,
    objects: ObjectContainer,
    config: SimulationConfig,
    key=None,
    stopping_condition=None,
    show_progress: bool = True,
    progress_callback: Callable[[int, int], None] | None = None,
) -> SimulationState:
    """``wrapper.py:14-63`` dispatch rule."""
    if config.gradient_config is None:
        return checkpointed_fdtd(arrays, objects, config, key, stopping_condition, show_progress, progress_callback)
    if stopping_condition is not None:
        raise Exception("Custom stopping conditions are only supported for forward-only simulations")
    if config.gradient_config.method == "reversible":
        return reversible_fdtd(arrays, objects, config, key, show_progress, progress_callback)
    if config.gradient_config.method == "checkpointed":
        return checkpointed_fdtd(arrays, objects, config, key, None, show_progress, progress_callback)
    raise Exception(f"Unknown gradient computation method: {config.gradient_config.method}")


def forward(state: SimulationState, config, objects, key=None, record_detectors=True, record_boundaries=False, simulate_boundaries=True) -> SimulationState:
    """One forward step (``forward.py:83-156``)."""
    t, arrays = state
    plan = get_plan(arrays, objects, config)
    plan.run_forward(int(t), 1, record_detectors, record_boundaries, simulate_boundaries)
    return int(t) + 1, plan.finish(arrays)


def backward(state: SimulationState, config, objects, key=None, record_detectors=True, reset_fields=True, fields_to_reset=("E", "H")) -> SimulationState:
    """One reverse step (``backward.py:62-135``)."""
    if tuple(fields_to_reset) != ("E", "H"):
        raise NotImplementedError("fields_to_reset other than ('E','H')")
    t, arrays = state
    plan = get_plan(arrays, objects, config)
    plan.run_reverse(int(t), 1, record_detectors, reset_fields)
    return int(t) - 1, plan.finish(arrays)


def full_backward(state: SimulationState, objects, config, key=None, record_detectors=True, reset_fields=True, start_time_step: int = 0) -> SimulationState:
    """``backward.py:18-59``."""
    t, arrays = state
    n = int(t) - int(start_time_step)
    if n <= 0:
        return state
    plan = get_plan(arrays, objects, config)
    plan.run_reverse(int(t), n, record_detectors, reset_fields)
    return int(start_time_step), plan.finish(arrays)


def apply_params(arrays, objects, params, key=None, **transform_kwargs):
    """``initialization.py:317-521``: latent device parameters -> ``inv_permittivities`` on the GPU,
    differentiably (``fdtdx_b200/device.py``).  With no devices the arrays pass through unchanged."""
    from fdtdx_b200.device import apply_params as _apply

    return _apply(arrays, objects, params or {}, key, **transform_kwargs)
