"""On/off schedules (reference ``fdtdx/core/switch.py:1-215``) and ``WaveCharacter``
(``fdtdx/core/wavelength.py:6-65``).  Host-side only: their outputs become the per-step gate
tables the kernels read."""

from __future__ import annotations

import math
from dataclasses import dataclass

from fdtdx_b200 import constants


@dataclass(frozen=True)
class WaveCharacter:
    phase_shift: float = 0.0
    period: float | None = None
    wavelength: float | None = None
    frequency: float | None = None

    def __post_init__(self):
        if sum(x is not None for x in (self.period, self.frequency, self.wavelength)) != 1:
            raise Exception("Need to set exactly one of Period, Frequency or Wavelength in WaveCharacter")

    def get_period(self) -> float:
        if self.period is not None:
            return self.period
        if self.wavelength is not None:
            return self.wavelength / constants.c
        return 1.0 / self.frequency

    def get_wavelength(self) -> float:
        if self.wavelength is not None:
            return self.wavelength
        if self.period is not None:
            return self.period * constants.c
        return constants.c / self.frequency

    def get_frequency(self) -> float:
        if self.frequency is not None:
            return self.frequency
        if self.period is not None:
            return 1.0 / self.period
        return constants.c / self.wavelength


@dataclass(frozen=True)
class OnOffSwitch:
    start_time: float | None = None
    start_after_periods: float | None = None
    end_time: float | None = None
    end_after_periods: float | None = None
    on_for_time: float | None = None
    on_for_periods: float | None = None
    period: float | None = None
    fixed_on_time_steps: tuple[int, ...] | list[int] | None = None
    is_always_off: bool = False
    interval: int = 1

    @property
    def is_default_always_on(self) -> bool:
        return (
            not self.is_always_off
            and self.fixed_on_time_steps is None
            and self.interval == 1
            and all(
                x is None
                for x in (
                    self.start_time,
                    self.start_after_periods,
                    self.end_time,
                    self.end_after_periods,
                    self.on_for_time,
                    self.on_for_periods,
                    self.period,
                )
            )
        )

    def is_on_at_time_step(self, time_step: int, time_step_duration: float) -> bool:
        """``switch.py:112-208``."""
        if self.is_always_off:
            return False
        start_time, end_time, on_for_time = self.start_time, self.end_time, self.on_for_time
        sap, eap, ofp, period = self.start_after_periods, self.end_after_periods, self.on_for_periods, self.period
        if any(x is not None for x in (sap, eap, ofp)) and period is None:
            raise Exception("Need to specify period!")
        num_start = sum(
            [
                start_time is not None,
                sap is not None,
                on_for_time is not None and end_time is not None,
                ofp is not None and end_time is not None,
                on_for_time is not None and eap is not None,
                ofp is not None and eap is not None,
            ]
        )
        if num_start > 1:
            raise Exception("Invalid start time specification!")
        if num_start == 0:
            start_time = 0
        num_end = sum(
            [
                end_time is not None,
                eap is not None,
                on_for_time is not None and start_time is not None,
                ofp is not None and start_time is not None,
                on_for_time is not None and sap is not None,
                ofp is not None and sap is not None,
            ]
        )
        if num_end > 1:
            raise Exception("Invalid end time specification!")
        if num_end == 0:
            end_time = math.inf
        if sap is not None:
            start_time = sap * period
        if eap is not None:
            end_time = eap * period
        if ofp is not None:
            on_for_time = ofp * period
        if start_time is None and on_for_time is not None:
            start_time = end_time - on_for_time
        if end_time is None and on_for_time is not None:
            end_time = start_time + on_for_time
        time_passed = time_step * time_step_duration
        return (start_time <= time_passed) and (time_passed <= end_time)

    def calculate_on_list(self, num_total_time_steps: int, time_step_duration: float) -> list[bool]:
        if self.fixed_on_time_steps is not None:
            on_list = [False] * num_total_time_steps
            for t in self.fixed_on_time_steps:
                on_list[t] = True
            return on_list
        return [
            self.is_on_at_time_step(t, time_step_duration) and t % self.interval == 0
            for t in range(num_total_time_steps)
        ]

    def calculate_time_step_to_on_arr_idx(self, num_total_time_steps: int, time_step_duration: float) -> list[int]:
        on_list = self.calculate_on_list(num_total_time_steps, time_step_duration)
        out, counter = [-1] * num_total_time_steps, 0
        for t in range(num_total_time_steps):
            if on_list[t]:
                out[t] = counter
                counter += 1
        return out
