"""Plan compiler: lowers ``(ObjectContainer, SimulationConfig, ArrayContainer)`` to the POD tables
the C ABI takes (``include/fdtdx_b200.h``) and binds the caller's device buffers.

What it reads off each object is exactly the hook list of SURVEY.md section 8b: boundary
axis/direction/slice/wrap flag and CPML tables; source arrays, profile and switch tables; detector
region, gate/index tables and reduction flags; recorder slot tables.
"""

from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from fdtdx_b200 import _lib
from fdtdx_b200._lib import check
from fdtdx_b200.boundaries import PerfectElectricConductor, PerfectMagneticConductor
from fdtdx_b200.constants import c as c0
from fdtdx_b200.detectors import (
    COMPONENT_NAMES,
    ClosedSurfacePhasorPoyntingFluxDetector,
    ClosedSurfacePoyntingFluxDetector,
    EnergyDetector,
    FieldDetector,
    PhasorDetector,
    PoyntingFluxDetector,
    phasor_table,
)
from fdtdx_b200.profile import PROFILE_CW, PROFILE_PULSE, PROFILE_TABLE
from fdtdx_b200.sources import PointDipoleSource, TFSFPlaneSource

_f32 = np.float32


def _fptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _iarr(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def metric_scales(config, axis: int):
    """(sB, sF) float32 per-cell scales of ``curl.py:10-39`` for one axis (global length)."""
    grid = config.resolved_grid
    w = grid.cell_widths(axis)
    ref = c0 * config.time_step_duration / config.courant_number
    prev = np.concatenate([w[:1], w[:-1]])
    wb = _f32(0.5) * (w + prev)
    return (_f32(ref) / wb).astype(_f32), (_f32(ref) / w).astype(_f32)


def clip_detector(det, x0: int, x1: int, config):
    """This rank's part of a detector whose region straddles an x-slab edge (sources,
    detectors and PML slabs are clipped to each rank's slab): a shallow copy over the clipped region with
    its weight tables recomputed; ``_global`` keeps the whole-region object (global normalisations,
    state merging in ``dist.merge_detector_states``).  Returns ``det`` itself when nothing is cut."""
    import copy

    dlo, dhi = det.grid_slice_tuple[0]
    lo, hi = max(dlo, x0), min(dhi, x1)
    if (lo, hi) == (dlo, dhi):
        return det
    if isinstance(det, (ClosedSurfacePhasorPoyntingFluxDetector, ClosedSurfacePoyntingFluxDetector)) or type(det).__name__ in ("ModeOverlapDetector", "PhasorPoyntingFluxDetector"):
        raise NotImplementedError(f"detector {det.name!r} ({type(det).__name__}) straddles the slab edge at x = {x0 if dlo < x0 else x1}")
    c = copy.copy(det)
    if isinstance(det, PoyntingFluxDetector):
        c.fixed_propagation_axis = det.propagation_axis  # not re-inferred from the clipped shape
    c.grid_slice_tuple = ((lo, hi), det.grid_slice_tuple[1], det.grid_slice_tuple[2])
    c.place_on_grid(config)
    if isinstance(det, EnergyDetector) and det._slice_indices is not None:
        gi = dlo + det._slice_indices[0]  # the fixed-x plane lives on exactly one rank
        c._slice_indices = (gi - lo if lo <= gi < hi else -1, det._slice_indices[1], det._slice_indices[2])
    c._global = det
    return c


def _profile_params(profile, wave_character):
    p = [0.0] * 8
    signal = None
    if profile.kind == PROFILE_CW:
        period = wave_character.get_period()
        p[0], p[1], p[2], p[3] = period, wave_character.phase_shift, profile.phase_shift, profile.num_startup_periods * period
    elif profile.kind == PROFILE_PULSE:
        sw = profile.spectral_width.get_frequency()
        fc = profile.center_wave.get_frequency()
        sigma_t = 1.0 / (2 * np.pi * sw)
        p[0], p[1], p[2], p[3], p[4] = 2 * np.pi * fc, wave_character.phase_shift, profile.center_wave.phase_shift, 6 * sigma_t, 2.0 * sigma_t**2
    elif profile.kind == PROFILE_TABLE:
        p[0], p[1], p[2], p[3] = profile.start_time, profile.time_step_duration, profile.outside_value, 1.0 if profile.interpolation == "nearest" else 0.0
        signal = np.ascontiguousarray(profile.signal, dtype=_f32)
    else:
        raise NotImplementedError(type(profile))
    return (C.c_double * 8)(*p), signal


class Plan:
    """Owns one ``FdtdxPlan*``.  ``x_range`` restricts the plan to an x-slab (multi-GPU)."""

    def __init__(self, objects, config, arrays, x_range: tuple[int, int] | None = None, halo=(False, False), bloch_role: str | None = None):
        self.lib = _lib.lib()
        # bloch_role "re" / "im": this plan is one of the two real systems of a complex (Bloch k != 0) run
        # (fdtdx_b200/bloch.py); the Im system carries no sources (a real injection enters the real part)
        self.bloch_role = bloch_role
        self.objects, self.config = objects, config
        shape = objects.volume.grid_shape
        self.global_shape = shape
        x0, x1 = x_range if x_range is not None else (0, shape[0])
        self.x0, self.x1 = x0, x1
        nx, ny, nz = x1 - x0, shape[1], shape[2]
        self.local_shape = (nx, ny, nz)
        self.T = max(int(config.time_steps_total), 1)
        self._keep = []  # host arrays must stay alive until the C call returns; kept for safety
        # Ragged rows (Nz % 4 != 0) cannot use 128-bit accesses or TMA tensor maps (16-byte strides).
        # For forward-only runs the plan then works on z-padded shadow copies: Nz is rounded up to a
        # multiple of 4, the extra cells stay at zero because their shadow inv_eps is zero (so they are the
        # zero halo the z-max face would see anyway), a z-max CPML slab is extended over them with zero
        # coefficients, and the caller's arrays are copied in before / out after every run call.
        self.pad = 0 if (x_range is not None or any(halo) or bloch_role is not None) else self._z_padding(objects, config, arrays, nz)
        self.nz_true = nz
        nz = nz + self.pad
        self._shadow = {}

        inv_eps, inv_mu = arrays.inv_permittivities, arrays.inv_permeabilities
        self.eps_tier = int(inv_eps.shape[0])
        self.mu_is_array = hasattr(inv_mu, "shape") and len(inv_mu.shape) > 0
        self.mu_tier = int(inv_mu.shape[0]) if self.mu_is_array else 0
        sE, sH = arrays.electric_conductivity, arrays.magnetic_conductivity
        self.sigE_tier = 0 if sE is None else int(sE.shape[0])
        self.sigH_tier = 0 if sH is None else int(sH.shape[0])
        wrap = [False, False, False]
        for b in objects.boundary_objects:
            if b.uses_wrap_padding:
                if getattr(b, "needs_complex_fields", False) and bloch_role is None:
                    raise ValueError("Bloch boundaries with k != 0 need complex fields: pass a container whose E / H are complex64")
                wrap[b.axis] = True
        # config.symmetry (the reduced half-domain itself is built at setup time, fdtd/symmetry.py - out of scope):
        # the step's part is the one-sided halo rule of update.py:121-125 and the detector mirror of :139-198
        self.sym = [int(s != 0) for s in config.symmetry]
        self.mirror = [int(config.symmetry[a] == -1 and any(getattr(b, "_is_symmetry_wall", False) and b.axis == a for b in objects.boundary_objects)) for a in range(3)]
        self.wrap = wrap
        sB = sF = widths = None
        if config.has_nonuniform_grid:
            sB_l, sF_l, w_l = [], [], []
            for a in range(3):
                b_, f_ = metric_scales(config, a)
                if a == 0:
                    b_, f_ = b_[x0:x1], f_[x0:x1]
                w_ = np.ascontiguousarray(config.resolved_grid.cell_widths(a), dtype=_f32)
                if a == 2 and self.pad:  # padded cells: any finite metric (they are masked to zero)
                    b_, f_, w_ = (np.concatenate([v, np.repeat(v[-1:], self.pad)]) for v in (b_, f_, w_))
                sB_l.append(np.ascontiguousarray(b_))
                sF_l.append(np.ascontiguousarray(f_))
                w_l.append(np.ascontiguousarray(w_, dtype=_f32))
            self._keep += sB_l + sF_l + w_l
            mk = lambda lst: (C.POINTER(C.c_float) * 3)(*[_fptr(x) for x in lst])
            sB, sF, widths = mk(sB_l), mk(sF_l), mk(w_l)
        h = C.c_void_p()
        check(
            self.lib.fdtdx_b200_plan_create(
                C.byref(h), nx, ny, nz, x0, shape[0], config.courant_number, config.time_step_duration, self.T,
                self.eps_tier, self.mu_tier, self.sigE_tier, self.sigH_tier,
                float(inv_mu) if not self.mu_is_array else 1.0, _iarr(wrap), sB, sF, widths,
            )
        )
        self.h = h
        if self.pad:
            check(self.lib.fdtdx_b200_set_z_padding(self.h, self.pad))
        if any(self.sym):
            if x_range is not None or bloch_role is not None:
                raise NotImplementedError("config.symmetry on x-sharded or complex (Bloch) plans")
            check(self.lib.fdtdx_b200_set_symmetry(self.h, _iarr(self.sym), _iarr(self.mirror)))
        self._add_boundaries()
        if bloch_role != "im":
            self._add_sources()
        self._add_detectors()
        if bloch_role is not None:
            if x_range is not None or any(halo):
                raise NotImplementedError("complex (Bloch) runs on x-sharded plans")
            cs, sn = [1.0, 1.0, 1.0], [0.0, 0.0, 0.0]
            for b in objects.boundary_objects:
                if getattr(b, "needs_complex_fields", False):
                    ph = b.get_bloch_phase(shape, config)
                    cs[b.axis], sn[b.axis] = float(np.real(ph)), float(np.imag(ph)) * (1.0 if bloch_role == "re" else -1.0)
            check(self.lib.fdtdx_b200_set_bloch(self.h, 1, (C.c_double * 3)(*cs), (C.c_double * 3)(*sn)))
        self._set_recorder()
        self.n_poles = 0
        if arrays.dispersive_c1 is not None:
            c1 = arrays.dispersive_c1
            self.n_poles = int(c1.shape[0])
            tiers = {int(x.shape[1]) for x in (arrays.dispersive_c1, arrays.dispersive_c2, arrays.dispersive_c3) if x is not None}
            if arrays.dispersive_c4 is not None:
                tiers.add(int(arrays.dispersive_c4.shape[1]))
            if len(tiers) != 1 or next(iter(tiers)) not in (1, 3):
                raise NotImplementedError(f"dispersive coefficient component tiers {tiers} (need one common tier of 1 or 3)")
            check(self.lib.fdtdx_b200_plan_set_dispersion(self.h, self.n_poles, next(iter(tiers)), int(arrays.dispersive_c4 is not None)))
        check(self.lib.fdtdx_b200_halo_bind(self.h, int(halo[0]), int(halo[1])))
        self.halo = halo
        self._bound = []

    @staticmethod
    def _z_padding(objects, config, arrays, nz: int) -> int:
        import os

        if nz % 4 == 0 or os.environ.get("FDTDX_B200_PAD_Z", "1") == "0":
            return 0
        gc = config.gradient_config
        if gc is not None and gc.method == "checkpointed":
            # that driver steps one call at a time; the shadow copies around every call would dominate
            return 0
        if os.environ.get("FDTDX_B200_PAD_Z_GRAD", "1") == "0" and (gc is not None or arrays.recording_state is not None):
            return 0
        if arrays.dispersive_c1 is not None or arrays.fields.dispersive_P_curr is not None:
            return 0
        tiers = [int(arrays.inv_permittivities.shape[0])]
        for a in (arrays.inv_permeabilities, arrays.electric_conductivity, arrays.magnetic_conductivity):
            if a is not None and hasattr(a, "shape") and len(a.shape) > 0:
                tiers.append(int(a.shape[0]))
        if any(t == 9 for t in tiers):
            return 0
        if not getattr(arrays.fields.E, "is_cuda", False):
            return 0
        for b in objects.boundary_objects:
            if b.uses_wrap_padding and b.axis == 2:
                return 0  # a periodic z axis wraps at the true Nz
        return (-nz) % 4

    # ------------------------------------------------------------------ tables
    def _add_boundaries(self):
        self.pml_index = {}
        self._zext = {}  # z-padded layout: cells added before / after each z slab's psi rows (pre, post)
        order = [p.axis for p in self.objects.pml_objects]
        if order != sorted(order):
            raise NotImplementedError("PML objects must be listed in axis order (x, y, z) - the CPML corrections are applied in that order")
        for pml in self.objects.pml_objects:
            lo, hi = pml.grid_slice_tuple[pml.axis]
            full = [(0, n) for n in self.global_shape]
            full[pml.axis] = (lo, hi)
            if tuple(full) != tuple(pml.grid_slice_tuple):
                raise NotImplementedError("PML slabs must span the full cross-section")
            t = [np.ascontiguousarray(x.reshape(-1), dtype=_f32) for x in (pml.pml_a_E, pml.pml_b_E, pml.inv_kappa_E, pml.pml_a_H, pml.pml_b_H, pml.inv_kappa_H)]
            for i in (2, 5):
                if t[i].size == 1:
                    t[i] = np.full(hi - lo, t[i][0], _f32)
            lo_true, hi_true = lo, hi
            if self.pad and pml.axis == 2:
                # z slabs of the padded layout are registered as supersets whose extra cells have a = b = 0 and
                # 1/kappa = 1 (psi stays 0, the correction is an exact zero): the z-max slab swallows the padded
                # cells, and a slab whose inner edge sits on an odd cell grows by one interior cell so that the
                # kernels can move its psi as aligned 64-bit halves (the unmasked CPML path of the staged kernels)
                pre, post = (lo % 2, self.pad) if pml.direction == "+" else (0, hi % 2)
                ext = lambda x, one: np.ascontiguousarray(np.concatenate([np.full(pre, one, _f32), x, np.full(post, one, _f32)]))
                t = [ext(x, 1.0 if i in (2, 5) else 0.0) for i, x in enumerate(t)]
                lo, hi = lo - pre, hi + post
                self._zext[pml.name] = (pre, post)
            idx = check(
                self.lib.fdtdx_b200_plan_add_pml(
                    self.h, pml.axis, 1 if pml.direction == "+" else 0, lo, hi, *[_fptr(x) for x in t], int(pml.kappa_is_one)
                )
            )
            if (lo, hi) != (lo_true, hi_true):
                check(self.lib.fdtdx_b200_plan_pml_set_true_range(self.h, idx, lo_true, min(hi_true, hi)))
            self.pml_index[pml.name] = idx
        for b in self.objects.boundary_objects:
            if isinstance(b, (PerfectElectricConductor, PerfectMagneticConductor)):
                kind = 0 if isinstance(b, PerfectElectricConductor) else 1
                lo = [s[0] for s in b.grid_slice_tuple]
                hi = [s[1] for s in b.grid_slice_tuple]
                check(self.lib.fdtdx_b200_plan_add_wall(self.h, kind, b.axis, _iarr(lo), _iarr(hi)))
        # z-padded layout: the padded cells stay at zero without any mask - their shadow inv_eps is 0 (bind), so the E
        # update adds c * K * 0 to a zero field, and with E = 0 there (and beyond, the zero halo) every difference the H
        # update of a padded cell takes vanishes.  They are therefore exactly the zero halo of the z-max face.

    def _switch_tables(self, src):
        if src.uses_default_switch:
            return None, None
        on = np.ascontiguousarray(src._is_on_at_time_step_arr, dtype=np.uint8)
        t_adj = np.ascontiguousarray(src._time_step_to_on_idx, dtype=_f32)
        on, t_adj = self._pad_T(on), self._pad_T(t_adj)
        self._keep += [on, t_adj]
        return on.ctypes.data_as(C.POINTER(C.c_uint8)), _fptr(t_adj)

    def _pad_T(self, a):
        if a.shape[0] >= self.T:
            return np.ascontiguousarray(a[: self.T])
        return np.ascontiguousarray(np.concatenate([a, np.zeros(self.T - a.shape[0], a.dtype)]))

    def _add_sources(self):
        cfg = self.config
        for src in self.objects.sources:
            params, signal = _profile_params(src.temporal_profile, src.wave_character)
            on, t_adj = self._switch_tables(src)
            sl = [list(s) for s in src.grid_slice_tuple]
            if isinstance(src, TFSFPlaneSource):
                if self.eps_tier == 9 or self.mu_tier == 9:
                    pass  # tensor rows are handled by the tensor kernels with the same descriptor
                xs = slice(max(sl[0][0], self.x0) - sl[0][0], min(sl[0][1], self.x1) - sl[0][0])
                if xs.stop <= xs.start:
                    continue
                cplx = np.iscomplexobj(src._E) or np.iscomplexobj(src._H)
                arrs = [np.ascontiguousarray(np.real(a)[:, xs], dtype=_f32) for a in (src._E, src._H, src._time_offset_E, src._time_offset_H)]
                lo = [max(sl[0][0], self.x0), sl[1][0], sl[2][0]]
                hi = [min(sl[0][1], self.x1), sl[1][1], sl[2][1]]
                cE = float(_f32(cfg.courant_number) * _f32(src.metric_scale_at_plane(cfg, "backward")))
                cH = float(_f32(cfg.courant_number) * _f32(src.metric_scale_at_plane(cfg, "forward")))
                hf = None if src._temporal_H_filter is None else np.ascontiguousarray(src._temporal_H_filter, dtype=_f32)
                si = check(
                    self.lib.fdtdx_b200_plan_add_plane_source(
                        self.h, _iarr(lo), _iarr(hi), src.propagation_axis, 1 if src.direction == "+" else -1,
                        *[_fptr(a) for a in arrs], src.temporal_profile.kind, params, _fptr(signal),
                        0 if signal is None else signal.shape[0], float(src.static_amplitude_factor), cE, cH, on, t_adj,
                        _fptr(hf), 0 if hf is None else hf.shape[0],
                    )
                )
                if cplx:
                    # lossy-mode profile: imaginary parts injected in quadrature (tfsf.py:266-283, 366-383)
                    im = [np.ascontiguousarray(np.imag(a)[:, xs], dtype=_f32) for a in (src._E, src._H)]
                    check(self.lib.fdtdx_b200_plan_source_set_quadrature(self.h, si, _fptr(im[0]), _fptr(im[1]), float(src.wave_character.phase_shift - 0.5 * np.pi)))
            elif isinstance(src, PointDipoleSource):
                cell = [s[0] for s in sl]
                if not (self.x0 <= cell[0] < self.x1):
                    continue
                scale = cfg.courant_number * src.amplitude * src.static_amplitude_factor
                check(
                    self.lib.fdtdx_b200_plan_add_dipole(
                        self.h, _iarr(cell), src.polarization, int(src.source_type == "electric"), scale,
                        src.temporal_profile.kind, params, _fptr(signal), 0 if signal is None else signal.shape[0], on, t_adj,
                    )
                )
            else:
                raise NotImplementedError(f"source type {type(src).__name__} is not on the hot path")

    def _add_detectors(self):
        cfg = self.config
        self.det_index, self.det_faces = {}, {}
        for det in self.objects.detectors:
            dlo, dhi = det.grid_slice_tuple[0]
            if dhi <= self.x0 or dlo >= self.x1:
                continue  # lives on another rank's slab
            if isinstance(det, ClosedSurfacePhasorPoyntingFluxDetector):
                # hollow shell (poynting_flux.py:391-503): one six-component phasor plane per face
                if self.x0 != 0 or self.x1 != self.global_shape[0]:
                    raise NotImplementedError("closed-surface phasor detectors on x-sharded plans")
                faces = []
                for key, sl in det.face_slices():
                    face = PhasorDetector(name=f"{det.name}/{key}", grid_slice_tuple=sl, wave_characters=det.wave_characters, components=COMPONENT_NAMES,
                                          scaling_mode=det.scaling_mode, dft_subsample=det.dft_subsample, switch=det.switch,
                                          exact_interpolation=det.exact_interpolation, inverse=det.inverse)
                    face.place_on_grid(cfg)
                    faces.append((key, self._add_detector(face)))
                self.det_faces[det.name] = faces
                continue
            local = clip_detector(det, self.x0, self.x1, cfg)
            di = self._add_detector(local)
            self.det_index[det.name] = di
            if local is not det and getattr(det, "reduce_volume", False) and isinstance(det, (FieldDetector, PhasorDetector)):
                # weighted mean over the WHOLE region: every rank divides its partial sum by the global weight sum
                check(self.lib.fdtdx_b200_plan_detector_set_wsum(self.h, di, float(np.sum(det._cached_cell_volume_weights, dtype=np.float64))))

    def _add_detector(self, det) -> int:
        cfg = self.config
        if True:
            flags = 0
            if det.exact_interpolation:
                flags |= _lib.DETF_EXACT
            if det.inverse:
                flags |= _lib.DETF_INVERSE
            comp_mask, aux, weights, nf, table, window, scale, slice_idx = 0, 0, None, 0, None, None, 1.0, [0, 0, 0]
            if isinstance(det, PhasorDetector):
                kind = _lib.DET_PHASOR
                comp_mask = sum(1 << i for i, n in enumerate(COMPONENT_NAMES) if n in det.components)
                if det.reduce_volume:
                    flags |= _lib.DETF_REDUCE
                    weights = det._cached_cell_volume_weights
                om = det._angular_frequencies
                nf = om.shape[0]
                table = phasor_table(det, self.T, cfg.time_step_duration)
                window = self._pad_T(np.ascontiguousarray(det._window_at_time_step_arr, dtype=_f32))
                scale = float(det._static_scale())
            elif isinstance(det, FieldDetector):
                kind = _lib.DET_FIELD
                comp_mask = sum(1 << i for i, n in enumerate(COMPONENT_NAMES) if n in det.components)
                if det.reduce_volume:
                    flags |= _lib.DETF_REDUCE
                    weights = det._cached_cell_volume_weights
            elif isinstance(det, EnergyDetector):
                kind = _lib.DET_ENERGY
                if det.as_slices:
                    flags |= _lib.DETF_SLICES
                    if det.use_mean:
                        flags |= _lib.DETF_SLICE_MEAN
                    slice_idx = list(det._slice_indices)
                elif det.reduce_volume:
                    flags |= _lib.DETF_REDUCE
                    weights = det._cached_cell_volume_weights
            elif isinstance(det, ClosedSurfacePoyntingFluxDetector):
                # net flux through the box faces (poynting_flux.py:199-284): a staged per-cell sum of
                # (+/-) S_a * area over the face cells of every active axis, reduced like reduce_volume
                kind = _lib.DET_POYNTING
                flags |= _lib.DETF_CLOSED | _lib.DETF_REDUCE | _lib.DETF_KEEP_ALL
                aux = sum(1 << a for a in det._resolve_active_axes())
                if det.orientation == "inward":
                    flags |= _lib.DETF_NEGATIVE
                weights = det._face_area_weights_per_axis
            elif isinstance(det, PoyntingFluxDetector):
                kind = _lib.DET_POYNTING
                if det.keep_all_components:
                    flags |= _lib.DETF_KEEP_ALL
                else:
                    aux = det.propagation_axis
                if det.direction == "-":
                    flags |= _lib.DETF_NEGATIVE
                if det.reduce_volume:
                    flags |= _lib.DETF_REDUCE
                    weights = det._cached_face_area_weights
            else:
                raise NotImplementedError(f"detector type {type(det).__name__} is not on the hot path")
            lo = [s[0] for s in det.grid_slice_tuple]
            hi = [s[1] for s in det.grid_slice_tuple]
            # large exact regions (videos, volume reductions, whole cross-sections): row-marching kernels
            # with 128-bit accesses instead of one thread per cell (csrc/det_volume.cuh)
            ext = [h_ - l_ for l_, h_ in zip(lo, hi)]
            if det.exact_interpolation and ext[2] >= 32 and ext[0] * ext[1] * ext[2] >= 4096 and os.environ.get("FDTDX_B200_DET_VOLUME", "1") != "0" and self.bloch_role is None and not any(self.sym):
                surface_only = (isinstance(det, EnergyDetector) and det.as_slices and not det.use_mean) or isinstance(det, ClosedSurfacePoyntingFluxDetector)
                if not surface_only:  # three planes / the box shell only: O(surface) work already
                    flags |= _lib.DETF_VOLUME
            on = self._pad_T(np.ascontiguousarray(det._is_on_at_time_step_arr, dtype=np.uint8))
            idx = self._pad_T(np.ascontiguousarray(det._time_step_to_arr_idx, dtype=np.int32))
            if isinstance(det, PhasorDetector):
                idx = np.zeros_like(idx)
            w = None if weights is None else np.ascontiguousarray(weights, dtype=_f32)
            di = check(
                self.lib.fdtdx_b200_plan_add_detector(
                    self.h, kind, _iarr(lo), _iarr(hi), flags, comp_mask, aux,
                    on.ctypes.data_as(C.POINTER(C.c_uint8)), idx.ctypes.data_as(C.POINTER(C.c_int32)),
                    _fptr(w), nf, _fptr(None if table is None else table.reshape(-1)), _fptr(window), scale, _iarr(slice_idx),
                )
            )
            return di

    def _set_recorder(self):
        gc = self.config.gradient_config
        self.recorder = None
        if gc is None or gc.recorder is None:
            return
        rec = gc.recorder
        if rec.slot_of_time is None or rec._max_time_steps != self.config.time_steps_total:
            rec.init_tables(self.config.time_steps_total)
        self.recorder = rec
        i32 = lambda a: self._pad_T(np.ascontiguousarray(a, dtype=np.int32)).ctypes.data_as(C.POINTER(C.c_int32))
        w = self._pad_T(np.ascontiguousarray(rec.replay_w, dtype=_f32))
        check(
            self.lib.fdtdx_b200_plan_set_recorder(
                self.h, rec.dtype_code, rec._latent_array_size, i32(rec.slot_of_time), i32(rec.replay_a), i32(rec.replay_b), _fptr(w)
            )
        )

    # ------------------------------------------------------------------ binding
    def _bind(self, slot: int, index: int, t, dtype=None, shape=None):
        import torch

        if t is None:
            check(self.lib.fdtdx_b200_bind(self.h, slot, index, None))
            return
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("fdtdx_b200 buffers must be contiguous CUDA tensors")
        if dtype is not None and t.dtype != dtype:
            raise ValueError(f"buffer dtype {t.dtype} != expected {dtype}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"buffer shape {tuple(t.shape)} != expected {tuple(shape)}")
        self._bound.append(t)
        check(self.lib.fdtdx_b200_bind(self.h, slot, index, C.c_void_p(t.data_ptr())))

    def _bind_z(self, slot: int, index: int, t, dtype, shape, state: bool, fill: float = 0.0, pre: int = 0, post: int | None = None):
        """Bind an array whose last axis runs along z.  Without padding: the array itself.  With padding:
        a plan-owned shadow whose last axis is longer (``pre`` cells in front, ``post`` - default: the z padding -
        behind); ``state`` arrays are copied back after every run call, constants (materials) only copied in."""
        import torch

        if not self.pad:
            return self._bind(slot, index, t, dtype, shape)
        post = self.pad if post is None else post
        if pre == 0 and post == 0:
            return self._bind(slot, index, t, dtype, shape)
        if dtype is not None and t.dtype != dtype:
            raise ValueError(f"buffer dtype {t.dtype} != expected {dtype}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"buffer shape {tuple(t.shape)} != expected {tuple(shape)}")
        n = t.shape[-1]
        key = (slot, index)
        sh = self._shadow.get(key)
        want = (*t.shape[:-1], pre + n + post)
        if sh is None or tuple(sh[1].shape) != want or sh[1].device != t.device or sh[4] != pre:
            sh = [t, torch.full(want, fill, dtype=t.dtype, device=t.device), n, state, pre]
            self._shadow[key] = sh
        sh[0], sh[3] = t, state
        self._bind(slot, index, sh[1], dtype)

    def _sync_in(self):
        for t, buf, n, _, pre in self._shadow.values():
            buf[..., pre:pre + n].copy_(t)

    def _sync_out(self):
        for t, buf, n, state, pre in self._shadow.values():
            if state:
                t.copy_(buf[..., pre:pre + n])

    def bind(self, arrays):
        import torch

        f32 = torch.float32
        self._bound = []
        L = self.local_shape
        self._bind_z(_lib.SLOT_E, 0, arrays.fields.E, f32, (3, *L), True)
        self._bind_z(_lib.SLOT_H, 0, arrays.fields.H, f32, (3, *L), True)
        self._bind_z(_lib.SLOT_INV_EPS, 0, arrays.inv_permittivities, f32, (self.eps_tier, *L), False, fill=0.0)  # 0: padded cells never move
        if self.mu_is_array:
            self._bind_z(_lib.SLOT_INV_MU, 0, arrays.inv_permeabilities, f32, (self.mu_tier, *L), False, fill=1.0)
        if self.sigE_tier:
            self._bind_z(_lib.SLOT_SIGMA_E, 0, arrays.electric_conductivity, f32, (self.sigE_tier, *L), False)
        if self.sigH_tier:
            self._bind_z(_lib.SLOT_SIGMA_H, 0, arrays.magnetic_conductivity, f32, (self.sigH_tier, *L), False)
        for pml in self.objects.pml_objects:
            q = self.pml_index[pml.name]
            if pml.name not in arrays.fields.psi_E:
                continue  # slab lives on another rank
            # x / y slabs run along the padded z axis; z slabs are thicker by their (pre, post) extension
            pre, post = self._zext.get(pml.name, (0, None)) if pml.axis == 2 else (0, None)
            for w in range(2):
                if self.pad:
                    self._bind_z(_lib.SLOT_PSI_E, 2 * q + w, arrays.fields.psi_E[pml.name][w], f32, None, True, pre=pre, post=post)
                    self._bind_z(_lib.SLOT_PSI_H, 2 * q + w, arrays.fields.psi_H[pml.name][w], f32, None, True, pre=pre, post=post)
                else:
                    self._bind(_lib.SLOT_PSI_E, 2 * q + w, arrays.fields.psi_E[pml.name][w], f32)
                    self._bind(_lib.SLOT_PSI_H, 2 * q + w, arrays.fields.psi_H[pml.name][w], f32)
        if self.n_poles:
            self._bind(_lib.SLOT_P_A, 0, arrays.fields.dispersive_P_curr, f32, (self.n_poles, 3, *L))
            self._bind(_lib.SLOT_P_B, 0, arrays.fields.dispersive_P_prev, f32, (self.n_poles, 3, *L))
            self._bind(_lib.SLOT_C1, 0, arrays.dispersive_c1, f32)
            self._bind(_lib.SLOT_C2, 0, arrays.dispersive_c2, f32)
            self._bind(_lib.SLOT_C3, 0, arrays.dispersive_c3, f32)
            if arrays.dispersive_c4 is not None:
                self._bind(_lib.SLOT_C4, 0, arrays.dispersive_c4, f32)
            check(self.lib.fdtdx_b200_set_parity(self.h, 0, 0, 0))
        for det in self.objects.detectors:
            if det.name in self.det_faces:
                for key, di in self.det_faces[det.name]:
                    self._bind(_lib.SLOT_DET_STATE, 4 * di, arrays.detector_states[det.name][key])
                continue
            if det.name not in self.det_index:
                continue
            di = self.det_index[det.name]
            st = arrays.detector_states[det.name]
            for k, key in enumerate(st.keys()):
                self._bind(_lib.SLOT_DET_STATE, 4 * di + k, st[key])
        if self.recorder is not None and arrays.recording_state is not None:
            for pml in self.objects.pml_objects:
                q = self.pml_index[pml.name]
                for f, fs in enumerate(("E", "H")):
                    buf = arrays.recording_state.data.get(f"{pml.name}_{fs}")
                    if buf is None:
                        continue  # x-sharded: this x interface plane lives on another rank
                    self._bind(_lib.SLOT_REC_DATA, 2 * q + f, buf, self.recorder.torch_dtype())
        self._bind_tensor_path(arrays)

    def _bind_tensor_path(self, arrays):
        """Full-tensor tier: ping-pong buffers and precomputed A/B (``fdtd/misc.py:69-129``)."""
        import torch

        self.tensor_E = self.eps_tier == 9 or self.sigE_tier == 9
        self.tensor_H = self.mu_tier == 9 or self.sigH_tier == 9
        if not (self.tensor_E or self.tensor_H):
            return
        from fdtdx_b200.tensor_setup import update_matrices

        def stamp(*ts):
            # identity + in-place version of the material tensors: the 3x3 solves are redone only when
            # the materials change (per-step drivers call bind() every step)
            return tuple(None if t is None or not hasattr(t, "data_ptr") else (t.data_ptr(), t._version, tuple(t.shape)) for t in ts)

        def alt_like(name, ref):
            buf = getattr(self, name, None)
            if buf is None or buf.shape != ref.shape or buf.device != ref.device:
                buf = torch.zeros_like(ref)
                setattr(self, name, buf)
            return buf

        if self.tensor_E:
            self._bind(_lib.SLOT_E_ALT, 0, alt_like("E_alt", arrays.fields.E))
            key = stamp(arrays.inv_permittivities, arrays.electric_conductivity)
            if getattr(self, "_tE_key", None) != key:
                A, B, Ar, Br = update_matrices(arrays.inv_permittivities, arrays.electric_conductivity, self.config.courant_number, "E")
                if arrays.electric_conductivity is None:
                    A = Ar = None  # M1 = M2 = I: A is exactly the identity (fdtd/misc.py:88-96), not stored
                self.tE, self._tE_key = (A, B, Ar, Br), key
        if self.tensor_H:
            self._bind(_lib.SLOT_H_ALT, 0, alt_like("H_alt", arrays.fields.H))
            key = stamp(arrays.inv_permeabilities, arrays.magnetic_conductivity)
            if getattr(self, "_tH_key", None) != key:
                A, B, Ar, Br = update_matrices(arrays.inv_permeabilities, arrays.magnetic_conductivity, self.config.courant_number, "H")
                if arrays.magnetic_conductivity is None:
                    A = Ar = None
                self.tH, self._tH_key = (A, B, Ar, Br), key
        self.set_tensor_direction(reverse=False)
        check(self.lib.fdtdx_b200_set_parity(self.h, 0, 0, 0))

    def set_tensor_direction(self, reverse: bool):
        if getattr(self, "tensor_E", False):
            A, B, Ar, Br = self.tE
            self._bind(_lib.SLOT_TENSOR_A_E, 0, Ar if reverse else A)
            self._bind(_lib.SLOT_TENSOR_B_E, 0, Br if reverse else B)
        if getattr(self, "tensor_H", False):
            A, B, Ar, Br = self.tH
            self._bind(_lib.SLOT_TENSOR_A_H, 0, Ar if reverse else A)
            self._bind(_lib.SLOT_TENSOR_B_H, 0, Br if reverse else B)

    # ------------------------------------------------------------------ execution
    @staticmethod
    def _stream():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run_forward(self, t0: int, n: int, record_detectors: bool, record_boundaries: bool, simulate_boundaries: bool = True):
        self.set_tensor_direction(False)
        self._sync_in()
        check(self.lib.fdtdx_b200_run_forward(self.h, int(t0), int(n), int(record_detectors), int(record_boundaries), int(simulate_boundaries), self._stream()))
        self._sync_out()

    def run_forward_phase(self, t: int, phase: int, record_detectors: bool, record_boundaries: bool, simulate_boundaries: bool = True):
        self._sync_in()
        check(self.lib.fdtdx_b200_run_forward_phase(self.h, int(t), int(phase), int(record_detectors), int(record_boundaries), int(simulate_boundaries), self._stream()))
        self._sync_out()

    def run_reverse_phase(self, t: int, phase: int, record_detectors: bool, reset_fields: bool):
        check(self.lib.fdtdx_b200_run_reverse_phase(self.h, int(t), int(phase), int(record_detectors), int(reset_fields), self._stream()))

    def bind_bloch_partner(self, E, H):
        self._bind(_lib.SLOT_BLOCH_E, 0, E)
        self._bind(_lib.SLOT_BLOCH_H, 0, H)

    def run_reverse(self, t_from: int, n: int, record_detectors: bool, reset_fields: bool):
        self.set_tensor_direction(True)
        self._sync_in()
        check(self.lib.fdtdx_b200_run_reverse(self.h, int(t_from), int(n), int(record_detectors), int(reset_fields), self._stream()))
        self._sync_out()

    def run_adjoint(self, arrays, t_from: int, n: int, cot_E, cot_H, cot_det: dict, grad_inv_eps, grad_inv_mu=None, keep_cot_psi: bool = False, exact: bool = False,
                    cot_P=None, cot_P_prev=None, grad_coeffs=None):
        """``fdtd_bwd`` loop (``fdtd/fdtd.py:262-333``): n iterations of reverse step + VJP of one forward
        step.  ``arrays`` holds the state at ``t_from`` (fields are reconstructed in place); ``cot_E`` /
        ``cot_H`` carry the field cotangents in place; ``cot_det[name][key]`` are the detector-state
        cotangents; gradients accumulate into ``grad_inv_eps`` / ``grad_inv_mu``."""
        import torch

        if any(self.sym):
            raise NotImplementedError("gradients with config.symmetry (the adjoint kernels transpose the two-sided halo rule)")
        self.bind(arrays)
        self._bind_z(_lib.SLOT_COT_E, 0, cot_E, torch.float32, None, True)
        self._bind_z(_lib.SLOT_COT_H, 0, cot_H, torch.float32, None, True)
        self._bind_z(_lib.SLOT_GRAD_INV_EPS, 0, grad_inv_eps, torch.float32, None, True)
        if grad_inv_mu is None:
            self._bind(_lib.SLOT_GRAD_INV_MU, 0, None)
        else:
            self._bind_z(_lib.SLOT_GRAD_INV_MU, 0, grad_inv_mu, None, None, True)
        if self.n_poles:
            # ADE adjoint: cotangents of P / P_prev carried in place, optional coefficient gradients
            self._bind(_lib.SLOT_COT_P, 0, cot_P, torch.float32)
            self._bind(_lib.SLOT_COT_P_PREV, 0, cot_P_prev, torch.float32)
            for k, slot in enumerate((_lib.SLOT_GRAD_C1, _lib.SLOT_GRAD_C2, _lib.SLOT_GRAD_C3, _lib.SLOT_GRAD_C4)):
                self._bind(slot, 0, None if grad_coeffs is None else grad_coeffs[k])
        if not keep_cot_psi or not getattr(self, "_cot_psi", None):
            self._cot_psi = {}
        for pml in self.objects.pml_objects:
            q = self.pml_index[pml.name]
            for w in range(2):
                for slot, src in ((_lib.SLOT_COT_PSI_E, arrays.fields.psi_E), (_lib.SLOT_COT_PSI_H, arrays.fields.psi_H)):
                    key = (slot, q, w)
                    if key not in self._cot_psi:
                        # plan-owned cotangent of psi, in the layout the kernels see (z-padded where psi is)
                        shp = list(src[pml.name][w].shape)
                        if self.pad:
                            shp[-1] += sum(self._zext.get(pml.name, (0, 0))) if pml.axis == 2 else self.pad
                        self._cot_psi[key] = torch.zeros(shp, dtype=torch.float32, device=src[pml.name][w].device)
                    self._bind(slot, 2 * q + w, self._cot_psi[key])
        for det in self.objects.detectors:
            if det.name in self.det_faces:
                for key, di in self.det_faces[det.name]:
                    self._bind(_lib.SLOT_COT_DET, 4 * di, cot_det.get(det.name, {}).get(key))
                continue
            if det.name not in self.det_index:
                continue
            di = self.det_index[det.name]
            keys = list(arrays.detector_states[det.name].keys())
            for k in range(4):
                g = cot_det.get(det.name, {}).get(keys[k]) if k < len(keys) else None
                self._bind(_lib.SLOT_COT_DET, 4 * di + k, g)
        self._sync_in()
        if exact:
            assert n == 1
            check(self.lib.fdtdx_b200_run_adjoint_exact(self.h, int(t_from) - 1, self._stream()))
        else:
            check(self.lib.fdtdx_b200_run_adjoint(self.h, int(t_from), int(n), self._stream()))
        self._sync_out()

    def parity(self) -> tuple[int, int, int]:
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.fdtdx_b200_get_parity(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def xchunk_hint(self) -> int:
        return check(self.lib.fdtdx_b200_get_xchunk(self.h))

    def launch_count(self) -> int:
        return int(self.lib.fdtdx_b200_launch_count(self.h))

    def set_tuning(self, xchunk: int = 0, rows: int = 0):
        check(self.lib.fdtdx_b200_set_tuning(self.h, int(xchunk), int(rows)))

    def total_energy(self, arrays):
        """Sum over the grid of ``compute_energy`` (metrics.py:15-67) as a 1-element device tensor."""
        import torch

        self.bind(arrays)
        self._sync_in()
        if getattr(self, "_energy_out", None) is None:
            self._energy_out = torch.zeros(1, dtype=torch.float32, device=arrays.fields.E.device)
        check(self.lib.fdtdx_b200_total_energy(self.h, C.c_void_p(self._energy_out.data_ptr()), self._stream()))
        return self._energy_out

    def set_tma(self, enable: int = -1, xchunk_tma: int = 0):
        """Select the TMA-staged (1) or register-marching (0) half-step kernels; -1 follows FDTDX_B200_TMA."""
        check(self.lib.fdtdx_b200_set_tma(self.h, int(enable), int(xchunk_tma)))

    def finish(self, arrays):
        """Return the container whose leaves hold the *current* state after ping-pong passes."""
        pp, ep, hp = self.parity()
        if self.n_poles and pp:
            arrays = arrays.aset("fields->dispersive_P_curr", arrays.fields.dispersive_P_prev).aset(
                "fields->dispersive_P_prev", arrays.fields.dispersive_P_curr
            )
        if getattr(self, "tensor_E", False) and ep:
            arrays.fields.E.copy_(self.E_alt)
        if getattr(self, "tensor_H", False) and hp:
            arrays.fields.H.copy_(self.H_alt)
        check(self.lib.fdtdx_b200_set_parity(self.h, 0, 0, 0))
        if self.n_poles and pp:
            # the swapped container is re-bound on the next call; nothing else to do
            pass
        return arrays

    def close(self):
        if getattr(self, "h", None):
            self.lib.fdtdx_b200_plan_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
