"""Stopping conditions of the forward loop (mirror of ``fdtd/stop_conditions.py:14-332``).

Same classes, attributes, defaults, validation messages and truth tables as the reference; a
condition is called with ``(state, config, objects)`` and returns True while the simulation should
continue.  The per-step global quantity each one needs is produced on the device:

* ``EnergyThresholdCondition`` - the total field energy is one fused reduction over the bound
  E / H / material arrays (``fdtdx_b200_total_energy``, csrc/aux_kernels.cuh), started only once
  ``min_steps`` has passed; one 4-byte read-back per step decides whether the loop continues.
* ``DetectorConvergenceCondition`` - reads ``(prev_periods + 1) * spp`` values of one reduced
  detector (a few hundred floats) and compares the two spectra on the host.

The drivers (``fdtd.checkpointed_fdtd``) step in bulk (`run_forward`) while a condition cannot fire
(before ``min_steps``) and one step at a time afterwards, so the stop step is exactly the reference's.
"""

from __future__ import annotations

import copy

import numpy as np


class StoppingCondition:
    """``stop_conditions.py:15-47``."""

    def setup(self, state, config, objects):
        return self

    def _validate(self, state, config, objects) -> None:
        pass

    def earliest_stop(self, config) -> int:
        """First time step at which the condition can return False (drivers step in bulk before it)."""
        return 0

    def __call__(self, state, config, objects) -> bool:
        raise NotImplementedError()

    def _with(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new


class TimeStepCondition(StoppingCondition):
    """``stop_conditions.py:51-78``: continue while ``time_steps_total > curr_time_step``."""

    def setup(self, state, config, objects):
        self._validate(state, config, objects)
        return self

    def earliest_stop(self, config) -> int:
        return int(config.time_steps_total)

    def __call__(self, state, config, objects) -> bool:
        curr_time_step, _ = state
        return bool(config.time_steps_total > int(curr_time_step))


def _total_energy(arrays, objects, config) -> float:
    """sum(compute_energy(E, H, inv_eps, inv_mu)) - on the device for CUDA containers."""
    E = arrays.fields.E
    if isinstance(E, np.ndarray):
        raise RuntimeError(
            "fdtdx_b200 stopping conditions evaluate on the device; NumPy containers belong to the oracle "
            "(oracle.yee.evaluate_condition)"
        )
    from fdtdx_b200.fdtd import get_plan

    plan = get_plan(arrays, objects, config)
    return float(plan.total_energy(arrays).item())


class EnergyThresholdCondition(StoppingCondition):
    """``stop_conditions.py:81-147``."""

    def __init__(self, threshold: float = 1e-6, min_steps: int | None = None, max_steps: int | None = None):
        self.threshold, self.min_steps, self.max_steps = threshold, min_steps, max_steps

    def setup(self, state, config, objects):
        new = self._with(
            max_steps=config.time_steps_total if self.max_steps is None else self.max_steps,
            min_steps=round(config.time_steps_total * 0.1) if self.min_steps is None else self.min_steps,
        )
        new._validate(state, config, objects)
        return new

    def _validate(self, state, config, objects) -> None:
        if self.threshold <= 0:
            raise ValueError(f"Energy threshold must be positive, got {self.threshold}.")
        if self.min_steps is not None and self.min_steps < 0:
            raise ValueError(f"Minimum steps must be non-negative, got {self.min_steps}.")

    def earliest_stop(self, config) -> int:
        return int(min(self.min_steps, self.max_steps))

    def decide(self, curr_time_step: int, total_energy: float) -> bool:
        """The reference's truth table (``:139-147``) given the reduced energy."""
        time_condition = curr_time_step < self.max_steps
        min_steps_condition = curr_time_step < self.min_steps
        converged = total_energy < self.threshold
        return bool(time_condition and (min_steps_condition or not converged))

    def __call__(self, state, config, objects) -> bool:
        if self.max_steps is None or self.min_steps is None:
            raise RuntimeError("EnergyThresholdCondition.setup() must be called before use. ")
        curr_time_step, arrays = state
        t = int(curr_time_step)
        if t < self.min_steps or t >= self.max_steps:  # the energy cannot change the answer
            return self.decide(t, float("inf"))
        return self.decide(t, _total_energy(arrays, objects, config))


class DetectorConvergenceCondition(StoppingCondition):
    """``stop_conditions.py:150-332``."""

    def __init__(self, detector_name: str, wave_character, prev_periods: int = 4, threshold: float = 1e-6,
                 min_steps: int | None = None, max_steps: int | None = None):
        self.detector_name, self.wave_character = detector_name, wave_character
        self.prev_periods, self.threshold = prev_periods, threshold
        self.min_steps, self.max_steps = min_steps, max_steps
        self._spp = None

    def setup(self, state, config, objects):
        spp = round(self.wave_character.get_period() / config.time_step_duration)
        new = self._with(
            _spp=spp,
            max_steps=config.time_steps_total if self.max_steps is None else self.max_steps,
            min_steps=round((self.prev_periods + 1) * spp) if self.min_steps is None else self.min_steps,
        )
        new._validate(state, config, objects)
        return new

    def _validate(self, state, config, objects) -> None:
        if self._spp is None:
            raise RuntimeError("DetectorConvergenceCondition: _spp was not initialized. Run setup() first.")
        _, arrays = state
        if (self.prev_periods + 1) * self._spp > config.time_steps_total:
            raise ValueError(
                "Number of samples over which DetectorConvergenceCondition computes is "
                "greater than the number of time steps in the simulation. "
                "Increase the time over which the simulation runs in SimulationConfig, "
                "decrease prev_periods, or use a source with a shorter period."
            )
        if self.detector_name not in arrays.detector_states:
            available = tuple(arrays.detector_states.keys())
            raise KeyError(f"Detector '{self.detector_name}' not found. Available detectors: {available}")
        det_state = arrays.detector_states[self.detector_name]
        if all(k not in det_state for k in ("energy", "poynting_flux", "fields")):
            available = tuple(det_state.keys())
            raise KeyError(
                f"Chosen detector does not seem to be an EnergyDetector, PoyntingFluxDetector, FieldDetector.\n "
                f"Available keys: {available}"
            )
        readings = next(iter(det_state.values()))
        if readings.ndim != 2:
            raise ValueError(
                f"The selected detector must have reduce_volume=True. Therefore, "
                f"the DetectorState('{self.detector_name}') array must have two "
                f"dimensions; got ndim={readings.ndim}.\n"
            )
        if readings.shape[0] != config.time_steps_total:
            raise ValueError(
                f"The number of detector readings must be exactly the same as the number of time steps in the simulation. "
                f"Number of detector readings: {readings.shape[0]}, time steps: {config.time_steps_total}.\n"
            )
        if self.prev_periods < 1:
            raise ValueError(f"prev_periods must be >= 1; got {self.prev_periods}.")
        if self.threshold < 0:
            raise ValueError(f"Detector convergence threshold must be non-negative, got {self.threshold}.")
        if self.min_steps is None:
            raise RuntimeError("DetectorConvergenceCondition: min_steps was not initialized.")
        if self.min_steps is not None and self.min_steps < (self.prev_periods + 1) * self._spp:
            raise ValueError(
                "min_steps must be larger than the number of steps used to compute convergence, "
                f"got {self.min_steps}, need more than {(self.prev_periods + 1) * self._spp}. "
                "You can also decrease prev_periods to match min_steps, or you can leave min_steps unset, "
                "as a suitable default will be used."
            )

    def earliest_stop(self, config) -> int:
        return int(min(self.min_steps, config.time_steps_total))

    def window(self, curr_time_step: int, total: int) -> tuple[int, int]:
        """(start_ref, start_last) with the reference's clamping (``:304-309``)."""
        spp, pp = self._spp, self.prev_periods
        start_ref = int(np.clip(curr_time_step - (pp + 1) * spp, 0, total - pp * spp))
        start_last = int(np.clip(curr_time_step - spp, 0, total - spp))
        return start_ref, start_last

    def converged(self, readings_ref: np.ndarray, readings_last: np.ndarray) -> bool:
        """``:311-324`` on host copies of the two reading windows (float32 like the reference)."""
        spp = self._spp
        ref_mean = np.mean(np.asarray(readings_ref, np.float32).reshape(self.prev_periods, spp), axis=0, dtype=np.float32)
        fft_ref = np.fft.rfft(ref_mean, n=spp)
        fft_last = np.fft.rfft(np.asarray(readings_last, np.float32), n=spp)
        distance = np.linalg.norm(np.abs(fft_ref) - np.abs(fft_last))
        return bool(distance < self.threshold)

    def __call__(self, state, config, objects) -> bool:
        if self._spp is None or self.min_steps is None or self.max_steps is None:
            raise RuntimeError("DetectorConvergenceCondition.setup() must be called before use.")
        curr_time_step, arrays = state
        t = int(curr_time_step)
        total = int(config.time_steps_total)
        time_condition = t < total
        min_steps_condition = t >= self.min_steps
        if not min_steps_condition:
            return True
        readings = next(iter(arrays.detector_states[self.detector_name].values()))
        a, b = self.window(t, total)
        spp, pp = self._spp, self.prev_periods
        ref = readings[a : a + pp * spp, 0]
        last = readings[b : b + spp, 0]
        if not isinstance(ref, np.ndarray):
            ref, last = ref.detach().cpu().numpy(), last.detach().cpu().numpy()
        return bool(time_condition and not self.converged(ref, last))
