"""ctypes binding of the C ABI (``include/fdtdx_b200.h``).

The product path has no CPU fallback: if ``libfdtdx_b200.so`` is missing or fails to load, every
entry point raises ``RuntimeError`` (build it with ``python -c 'import __graft_entry__ as g; g.build()'``
or ``fdtdx_b200/csrc/build.sh``).
"""

from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

LIB_PATH = os.environ.get("FDTDX_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfdtdx_b200.so")

# enum FdtdxSlot
(
    SLOT_E, SLOT_H, SLOT_INV_EPS, SLOT_INV_MU, SLOT_SIGMA_E, SLOT_SIGMA_H, SLOT_PSI_E, SLOT_PSI_H,
    SLOT_P_A, SLOT_P_B, SLOT_C1, SLOT_C2, SLOT_C3, SLOT_C4, SLOT_DET_STATE, SLOT_REC_DATA,
    SLOT_E_ALT, SLOT_H_ALT, SLOT_TENSOR_A_E, SLOT_TENSOR_B_E, SLOT_TENSOR_A_H, SLOT_TENSOR_B_H,
    SLOT_HALO_H_LO, SLOT_HALO_E_HI, SLOT_GRAD_INV_EPS, SLOT_GRAD_INV_MU, SLOT_COT_E, SLOT_COT_H,
    SLOT_COT_PSI_E, SLOT_COT_PSI_H, SLOT_COT_DET, SLOT_COT_P, SLOT_COT_P_PREV, SLOT_GRAD_C1, SLOT_GRAD_C2, SLOT_GRAD_C3,
    SLOT_GRAD_C4, SLOT_BLOCH_E, SLOT_BLOCH_H, SLOT_DET_XLO_E, SLOT_DET_XLO_H, SLOT_DET_XLO_HPREV, SLOT_COUNT,
) = range(43)

DET_FIELD, DET_ENERGY, DET_POYNTING, DET_PHASOR = 0, 1, 2, 3
DETF_EXACT, DETF_INVERSE, DETF_REDUCE, DETF_SLICES, DETF_SLICE_MEAN, DETF_KEEP_ALL, DETF_NEGATIVE, DETF_VOLUME, DETF_CLOSED = 1, 2, 4, 8, 16, 32, 64, 128, 256

EXPORTS = [
    "fdtdx_b200_last_error", "fdtdx_b200_version", "fdtdx_b200_plan_create", "fdtdx_b200_plan_destroy",
    "fdtdx_b200_plan_add_pml", "fdtdx_b200_plan_add_wall", "fdtdx_b200_plan_add_plane_source",
    "fdtdx_b200_plan_add_dipole", "fdtdx_b200_plan_add_detector", "fdtdx_b200_plan_set_recorder",
    "fdtdx_b200_plan_set_dispersion", "fdtdx_b200_halo_bind", "fdtdx_b200_bind", "fdtdx_b200_run_forward",
    "fdtdx_b200_run_forward_phase", "fdtdx_b200_run_reverse", "fdtdx_b200_run_adjoint",
    "fdtdx_b200_get_parity", "fdtdx_b200_set_parity", "fdtdx_b200_launch_count", "fdtdx_b200_set_tuning", "fdtdx_b200_set_tma", "fdtdx_b200_peer_export", "fdtdx_b200_peer_attach", "fdtdx_b200_peer_detach", "fdtdx_b200_total_energy", "fdtdx_b200_run_adjoint_exact",
    "fdtdx_b200_run_forward_host", "fdtdx_b200_run_half_range", "fdtdx_b200_get_xchunk",
    "fdtdx_b200_plan_source_set_quadrature", "fdtdx_b200_peer_status", "fdtdx_b200_set_bloch", "fdtdx_b200_run_reverse_phase", "fdtdx_b200_plan_detector_set_wsum", "fdtdx_b200_set_symmetry", "fdtdx_b200_set_z_padding", "fdtdx_b200_plan_pml_set_true_range",
]

_p = C.c_void_p
_i = C.c_int
_d = C.c_double
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_dp = C.POINTER(C.c_double)
_fpp = C.POINTER(_fp)


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fdtdx_b200: CUDA extension not built ({LIB_PATH} missing). There is no CPU fallback; "
            "run `python -c 'import __graft_entry__ as g; g.build()'`."
        )
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise RuntimeError(f"fdtdx_b200: cannot load {LIB_PATH}: {e}") from e
    L.fdtdx_b200_last_error.restype = C.c_char_p
    L.fdtdx_b200_version.restype = _i
    L.fdtdx_b200_plan_create.argtypes = [C.POINTER(_p), _i, _i, _i, _i, _i, _d, _d, _i, _i, _i, _i, _i, _d, _ip, _fpp, _fpp, _fpp]
    L.fdtdx_b200_plan_destroy.argtypes = [_p]
    L.fdtdx_b200_plan_add_pml.argtypes = [_p, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _i]
    L.fdtdx_b200_plan_add_wall.argtypes = [_p, _i, _i, _ip, _ip]
    L.fdtdx_b200_plan_add_plane_source.argtypes = [_p, _ip, _ip, _i, _i, _fp, _fp, _fp, _fp, _i, _dp, _fp, _i, _d, _d, _d, _u8p, _fp, _fp, _i]
    L.fdtdx_b200_plan_source_set_quadrature.argtypes = [_p, _i, _fp, _fp, _d]
    L.fdtdx_b200_plan_add_dipole.argtypes = [_p, _ip, _i, _i, _d, _i, _dp, _fp, _i, _u8p, _fp]
    L.fdtdx_b200_plan_add_detector.argtypes = [_p, _i, _ip, _ip, _i, _i, _i, _u8p, _i32p, _fp, _i, _fp, _fp, _d, _ip]
    L.fdtdx_b200_plan_set_recorder.argtypes = [_p, _i, _i, _i32p, _i32p, _i32p, _fp]
    L.fdtdx_b200_plan_set_dispersion.argtypes = [_p, _i, _i, _i]
    L.fdtdx_b200_halo_bind.argtypes = [_p, _i, _i]
    L.fdtdx_b200_set_bloch.argtypes = [_p, _i, _dp, _dp]
    L.fdtdx_b200_set_symmetry.argtypes = [_p, _ip, _ip]
    L.fdtdx_b200_set_z_padding.argtypes = [_p, _i]
    L.fdtdx_b200_plan_pml_set_true_range.argtypes = [_p, _i, _i, _i]
    L.fdtdx_b200_plan_detector_set_wsum.argtypes = [_p, _i, _d]
    L.fdtdx_b200_run_reverse_phase.argtypes = [_p, _i, _i, _i, _i, _p]
    L.fdtdx_b200_bind.argtypes = [_p, _i, _i, _p]
    L.fdtdx_b200_run_forward.argtypes = [_p, _i, _i, _i, _i, _i, _p]
    L.fdtdx_b200_run_forward_phase.argtypes = [_p, _i, _i, _i, _i, _i, _p]
    L.fdtdx_b200_run_half_range.argtypes = [_p, _i, _i, _i, _i, _i, _p]
    L.fdtdx_b200_get_xchunk.argtypes = [_p]
    L.fdtdx_b200_run_reverse.argtypes = [_p, _i, _i, _i, _i, _p]
    L.fdtdx_b200_run_adjoint.argtypes = [_p, _i, _i, _p]
    L.fdtdx_b200_get_parity.argtypes = [_p, _ip, _ip, _ip]
    L.fdtdx_b200_set_parity.argtypes = [_p, _i, _i, _i]
    L.fdtdx_b200_launch_count.argtypes = [_p]
    L.fdtdx_b200_launch_count.restype = C.c_longlong
    L.fdtdx_b200_set_tuning.argtypes = [_p, _i, _i]
    L.fdtdx_b200_set_tma.argtypes = [_p, _i, _i]
    L.fdtdx_b200_total_energy.argtypes = [_p, _p, _p]
    L.fdtdx_b200_run_adjoint_exact.argtypes = [_p, _i, _p]
    L.fdtdx_b200_peer_detach.argtypes = [_p]
    L.fdtdx_b200_peer_status.argtypes = [_p]
    L.fdtdx_b200_peer_export.argtypes = [_p, _i, C.c_char_p, C.POINTER(C.c_longlong)]
    L.fdtdx_b200_peer_attach.argtypes = [_p, _i, C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong, _i]
    L.fdtdx_b200_run_forward_host.argtypes = [_p, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int or name not in ("fdtdx_b200_last_error", "fdtdx_b200_launch_count"):
            if name != "fdtdx_b200_last_error" and name != "fdtdx_b200_launch_count":
                fn.restype = _i
    return L


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().fdtdx_b200_last_error().decode()
        if "Dispersive time-reversible" in msg:
            raise NotImplementedError(msg)
        raise RuntimeError(f"fdtdx_b200 error {rc}: {msg}")
    return rc
