"""PML-interface recorder configuration (host side).

Mirrors ``fdtdx/interfaces/{recorder.py:20-199, modules.py:98-161, time_filter.py:139-257}``:
a ``Recorder`` is an ordered module list; supported modules are ``DtypeConversion`` (cast the
recorded planes to bf16 / fp16 / fp8) and ``LinearReconstructEveryK`` (store steps 0,K,2K,...,T-1
and linearly interpolate on replay).  Record/replay itself runs in CUDA
(``csrc/interface_kernels.cuh``); this module only resolves the slot tables and buffer layout.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

REC_F32, REC_BF16, REC_F16, REC_F8E4M3FNUZ, REC_F8E4M3FN, REC_F8E5M2 = 0, 1, 2, 3, 4, 5
_DTYPE_CODES = {
    "float32": REC_F32,
    "bfloat16": REC_BF16,
    "float16": REC_F16,
    "float8_e4m3fnuz": REC_F8E4M3FNUZ,
    "float8_e4m3fn": REC_F8E4M3FN,
    "float8_e5m2": REC_F8E5M2,
}
_DTYPE_BYTES = {REC_F32: 4, REC_BF16: 2, REC_F16: 2, REC_F8E4M3FNUZ: 1, REC_F8E4M3FN: 1, REC_F8E5M2: 1}


def _dtype_name(dtype) -> str:
    name = getattr(dtype, "__name__", None) or str(dtype)
    name = name.replace("torch.", "").replace("jnp.", "").replace("numpy.", "")
    if name not in _DTYPE_CODES:
        raise ValueError(f"unsupported recorder dtype {dtype!r}")
    return name


@dataclass
class DtypeConversion:
    """``modules.py:98-161``."""

    dtype: object = "float32"
    exclude_filter: Sequence[str] = ()

    @property
    def code(self) -> int:
        return _DTYPE_CODES[_dtype_name(self.dtype)]


@dataclass
class LinearReconstructEveryK:
    """``time_filter.py:139-257``."""

    k: int = 1
    start_recording_after: int = 0
    _save_time_steps: np.ndarray | None = None
    _time_to_arr_idx: np.ndarray | None = None
    _array_size: int = 0

    def init_shapes(self, time_steps_max: int) -> "LinearReconstructEveryK":
        steps = list(range(self.start_recording_after, time_steps_max, self.k))
        if steps[-1] != time_steps_max - 1:
            steps.append(time_steps_max - 1)
        self._save_time_steps = np.asarray(steps, dtype=np.int32)
        self._array_size = len(steps)
        # time -> index of the last saved slot at or before it (time_filter.py:178-190)
        ti = np.zeros(time_steps_max, dtype=np.int32)
        ti[self._save_time_steps] = np.arange(self._array_size, dtype=np.int32)
        for _ in range(self.k - 1):
            rolled = np.roll(ti, 1)
            ti = np.where(ti == 0, rolled, ti)
            ti[: self.k] = 0
        self._time_to_arr_idx = ti
        return self


@dataclass
class Recorder:
    """``recorder.py:20-199``.  After ``init_state`` the recorder knows, for every time step,
    which latent slot it writes (``slot_of_time``; -1 = not saved) and how to reconstruct
    (``replay_table``: slot a, slot b, weight)."""

    modules: Sequence[object] = ()
    _max_time_steps: int = 0
    _latent_array_size: int = 0
    dtype_code: int = REC_F32
    slot_of_time: np.ndarray | None = None
    replay_a: np.ndarray | None = None
    replay_b: np.ndarray | None = None
    replay_w: np.ndarray | None = None

    def init_tables(self, max_time_steps: int) -> "Recorder":
        tfs = [m for m in self.modules if isinstance(m, LinearReconstructEveryK)]
        dcs = [m for m in self.modules if isinstance(m, DtypeConversion)]
        if len(tfs) > 1 or len(dcs) > 1:
            raise NotImplementedError("at most one time filter and one dtype conversion are supported")
        for m in self.modules:
            if not isinstance(m, (LinearReconstructEveryK, DtypeConversion)):
                raise NotImplementedError(f"recorder module {type(m).__name__} is not on the hot path")
        if dcs and tfs and list(self.modules).index(dcs[0]) < list(self.modules).index(tfs[0]):
            raise NotImplementedError("DtypeConversion must come after the time filter (interpolation runs in float32)")
        self._max_time_steps = max_time_steps
        self.dtype_code = dcs[0].code if dcs else REC_F32
        T = max_time_steps
        if tfs:
            tf = tfs[0].init_shapes(T)
            self._latent_array_size = tf._array_size
            slot = np.full(T, -1, np.int32)
            slot[tf._save_time_steps] = np.arange(tf._array_size, dtype=np.int32)
            a = tf._time_to_arr_idx.astype(np.int32).copy()
            saved = slot >= 0
            # linear_reconstruct (time_filter.py:237-250): first index in _time_to_arr_idx equal to
            # arr_idx / arr_idx + 1 is the save time of that slot (index_1d_array = argmax of ==)
            first_idx = lambda v: int(np.argmax(tf._time_to_arr_idx == v))
            w = np.zeros(T, np.float32)
            b = a.copy()
            for t in range(T):
                if saved[t]:
                    a[t], b[t], w[t] = slot[t], slot[t], 0.0
                else:
                    prev_t, next_t = first_idx(a[t]), first_idx(a[t] + 1)
                    b[t] = a[t] + 1
                    w[t] = np.float32(t - prev_t) / np.float32(next_t - prev_t)
            self.slot_of_time, self.replay_a, self.replay_b, self.replay_w = slot, a, b, w
        else:
            self._latent_array_size = T
            ar = np.arange(T, dtype=np.int32)
            self.slot_of_time, self.replay_a, self.replay_b = ar, ar.copy(), ar.copy()
            self.replay_w = np.zeros(T, np.float32)
        return self

    @property
    def elem_bytes(self) -> int:
        return _DTYPE_BYTES[self.dtype_code]

    def torch_dtype(self):
        import torch

        return {
            REC_F32: torch.float32,
            REC_BF16: torch.bfloat16,
            REC_F16: torch.float16,
            REC_F8E4M3FNUZ: torch.float8_e4m3fnuz,
            REC_F8E4M3FN: torch.float8_e4m3fn,
            REC_F8E5M2: torch.float8_e5m2,
        }[self.dtype_code]
