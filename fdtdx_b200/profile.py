"""Temporal source profiles (reference ``fdtdx/objects/sources/profile.py:253-439`` and
``fdtdx/core/window.py:16-30``).

``get_amplitude`` here is the float32 NumPy restatement used on the host (for tests and table
building); the kernels evaluate the same expressions in-kernel with the same rounding sequence
(``csrc/yee_kernels.cuh: source_amplitude``).  ``kind``/``params()`` is what the plan compiler
ships to the device.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from fdtdx_b200.switch import WaveCharacter

PROFILE_CW = 0
PROFILE_PULSE = 1
PROFILE_TABLE = 2

_f32 = np.float32


def _real_exp_minus_i(phase: np.ndarray) -> np.ndarray:
    # jnp.real(jnp.exp(-1j * phase)) == cos(phase) for real phase
    return np.cos(phase.astype(_f32)).astype(_f32)


@dataclass(frozen=True)
class SingleFrequencyProfile:
    """``profile.py:253-273``: ``cos(2*pi*t/T + phase) * clip(t / (n*T), 0, 1)``."""

    phase_shift: float = math.pi
    num_startup_periods: int = 4

    kind = PROFILE_CW

    def get_amplitude(self, time, period: float, phase_shift: float = 0.0):
        time = np.asarray(time, dtype=_f32)
        # 2*pi is a weak python scalar: (2*pi*time) / period + (phase_shift + self.phase_shift)
        time_phase = _f32(2 * np.pi) * time / _f32(period) + _f32(phase_shift) + _f32(self.phase_shift)
        raw = _real_exp_minus_i(time_phase)
        startup_time = self.num_startup_periods * period
        factor = np.clip(time / _f32(startup_time), _f32(0.0), _f32(1.0))
        return (factor * raw).astype(_f32)


@dataclass(frozen=True)
class GaussianPulseProfile:
    """``profile.py:276-345``."""

    spectral_width: WaveCharacter = None  # type: ignore[assignment]
    center_wave: WaveCharacter = None  # type: ignore[assignment]

    kind = PROFILE_PULSE

    def __post_init__(self):
        if self.spectral_width.phase_shift != 0.0:
            raise ValueError("spectral_width should not have a phase_shift.")

    def get_amplitude(self, time, period: float, phase_shift: float = 0.0):
        del period
        time = np.asarray(time, dtype=_f32)
        sw = self.spectral_width.get_frequency()
        fc = self.center_wave.get_frequency()
        sigma_t = 1.0 / (2 * np.pi * sw)
        t0 = 6 * sigma_t
        d = time - _f32(t0)
        envelope = np.exp(-(d * d) / _f32(2.0 * sigma_t**2)).astype(_f32)
        carrier_phase = _f32(2 * np.pi * fc) * time + _f32(phase_shift) + _f32(self.center_wave.phase_shift)
        return (envelope * _real_exp_minus_i(carrier_phase)).astype(_f32)


@dataclass(frozen=True)
class CustomTimeSignalProfile:
    """``profile.py:348-439``: sampled waveform with linear / nearest interpolation."""

    signal: np.ndarray = None  # type: ignore[assignment]
    time_step_duration: float = 0.0
    start_time: float = 0.0
    interpolation: str = "linear"
    outside_value: float = 0.0

    kind = PROFILE_TABLE

    def __post_init__(self):
        object.__setattr__(self, "signal", np.asarray(self.signal, dtype=_f32))
        if self.signal.ndim != 1:
            raise ValueError(f"signal must be one-dimensional, got shape {self.signal.shape}")
        if self.signal.shape[0] < 2:
            raise ValueError("signal must contain at least two samples")
        if self.time_step_duration <= 0:
            raise ValueError("time_step_duration must be positive")
        if self.interpolation not in ("linear", "nearest"):
            raise ValueError(f"interpolation must be 'linear' or 'nearest', got {self.interpolation!r}")

    def get_amplitude(self, time, period: float, phase_shift: float = 0.0):
        del period, phase_shift
        time = np.asarray(time, dtype=_f32)
        idx = (time - _f32(self.start_time)) / _f32(self.time_step_duration)
        floor_idx = np.floor(idx)
        idx0 = floor_idx.astype(np.int32)
        frac = idx - floor_idx
        n = self.signal.shape[0]
        valid = (idx0 >= 0) & (idx0 < n)
        i0 = np.clip(idx0, 0, n - 1)
        i1 = np.clip(i0 + 1, 0, n - 1)
        y0, y1 = self.signal[i0], self.signal[i1]
        if self.interpolation == "nearest":
            y = np.where(frac < 0.5, y0, y1)
        else:
            y = (_f32(1.0) - frac) * y0 + frac * y1
        return np.where(valid, y, _f32(self.outside_value)).astype(_f32)
