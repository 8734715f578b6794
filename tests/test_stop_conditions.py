"""Stopping conditions.  CPU part: the reference's own unit expectations
(tests/unit/fdtd/test_stop_conditions.py) re-typed against the host mirror classes (set-up,
validation, defaults) and against the oracle's restatement of the truth tables; GPU part: the
device energy reduction and the stop step / final state of a pulsed run vs the oracle."""

import numpy as np
import pytest

import fdtdx_b200 as fx
from fdtdx_b200.stop_conditions import DetectorConvergenceCondition, EnergyThresholdCondition, TimeStepCondition
from oracle import yee
from scenes import build_scene, rel_l2

F = np.float32
_CONFIG = fx.SimulationConfig(time=100e-11, grid=fx.UniformGrid(spacing=1e-3), courant_factor=0.99)


def _arrays(E=1.0, H=1.0, detector_states=None):
    return fx.ArrayContainer(
        fields=fx.FieldState(E=np.full((3, 4, 4, 4), E, F), H=np.full((3, 4, 4, 4), H, F), psi_E={}, psi_H={}),
        inv_permittivities=np.ones((1, 4, 4, 4), F), inv_permeabilities=np.ones((1, 4, 4, 4), F),
        detector_states=detector_states or {}, recording_state=None)


def _ev(cond, t, arrays=None):
    return yee.evaluate_condition(cond, (t, arrays if arrays is not None else _arrays()), _CONFIG, None)


# ---- TimeStepCondition (ref. TestTimeStepCondition) -------------------------------------------------
def test_time_step_condition_truth_table():
    cond = TimeStepCondition().setup((0, _arrays()), _CONFIG, None)
    T = _CONFIG.time_steps_total
    assert _ev(cond, T - 1) is True and _ev(cond, T) is False and _ev(cond, T + 5) is False
    assert cond((T - 1, None), _CONFIG, None) is True and cond((T, None), _CONFIG, None) is False


# ---- EnergyThresholdCondition (ref. TestEnergyThresholdCondition) -----------------------------------
def test_energy_validation_and_defaults():
    with pytest.raises(ValueError, match="positive"):
        EnergyThresholdCondition(threshold=0.0).setup((0, _arrays()), _CONFIG, None)
    with pytest.raises(ValueError, match="positive"):
        EnergyThresholdCondition(threshold=-1.0).setup((0, _arrays()), _CONFIG, None)
    with pytest.raises(ValueError, match="non-negative"):
        EnergyThresholdCondition(min_steps=-1).setup((0, _arrays()), _CONFIG, None)
    cond = EnergyThresholdCondition().setup((0, _arrays()), _CONFIG, None)
    assert cond.max_steps == _CONFIG.time_steps_total
    assert cond.min_steps == round(_CONFIG.time_steps_total * 0.1)
    assert EnergyThresholdCondition(max_steps=100).setup((0, _arrays()), _CONFIG, None).max_steps == 100
    assert EnergyThresholdCondition(min_steps=7).setup((0, _arrays()), _CONFIG, None).min_steps == 7
    with pytest.raises(RuntimeError, match="setup"):
        EnergyThresholdCondition()((0, _arrays()), _CONFIG, None)


def test_energy_truth_table():
    zero = _arrays(0.0, 0.0)
    cond = EnergyThresholdCondition(threshold=1e10, min_steps=20).setup((0, zero), _CONFIG, None)
    assert _ev(cond, 15, zero) is True  # below min_steps: continue although the energy is below the threshold
    cond = EnergyThresholdCondition(threshold=1e-100, min_steps=5).setup((0, _arrays()), _CONFIG, None)
    assert _ev(cond, 6) is True  # energy above threshold
    tiny = _arrays(1e-10, 1e-10)
    cond = EnergyThresholdCondition(threshold=1.0, min_steps=5).setup((0, tiny), _CONFIG, None)
    assert _ev(cond, 6, tiny) is False
    cond = EnergyThresholdCondition(threshold=1e-100, min_steps=5, max_steps=50).setup((0, _arrays()), _CONFIG, None)
    assert _ev(cond, 50) is False
    # the host mirror's decide() is the same table
    assert cond.decide(6, 1.0) is True and cond.decide(50, 1.0) is False and cond.decide(3, 0.0) is True
    assert EnergyThresholdCondition(threshold=1.0, min_steps=5, max_steps=50).decide(6, 0.5) is False


def test_oracle_energy_matches_closed_form():
    a = _arrays(2.0, 3.0)
    e = yee.compute_energy(a.fields.E, a.fields.H, a.inv_permittivities, a.inv_permeabilities)
    assert np.allclose(e, 0.5 * 3 * 4.0 + 0.5 * 3 * 9.0)


# ---- DetectorConvergenceCondition (ref. TestDetectorConvergenceCondition) ---------------------------
def _wc_for_spp(spp):
    return fx.WaveCharacter(period=spp * _CONFIG.time_step_duration)


def _det_arrays(readings, key="energy"):
    return _arrays(detector_states={"det": {key: readings}})


def test_detector_convergence_setup_and_validation():
    T = _CONFIG.time_steps_total
    ok = _det_arrays(np.zeros((T, 1), F))
    cond = DetectorConvergenceCondition("det", _wc_for_spp(10), prev_periods=4).setup((0, ok), _CONFIG, None)
    assert cond._spp == 10 and cond.max_steps == T and cond.min_steps == 50
    assert DetectorConvergenceCondition("det", _wc_for_spp(10), max_steps=77).setup((0, ok), _CONFIG, None).max_steps == 77
    with pytest.raises(KeyError, match="not found"):
        DetectorConvergenceCondition("nope", _wc_for_spp(10)).setup((0, ok), _CONFIG, None)
    with pytest.raises(KeyError, match="does not seem"):
        DetectorConvergenceCondition("det", _wc_for_spp(10)).setup((0, _det_arrays(np.zeros((T, 1), F), key="phasor")), _CONFIG, None)
    with pytest.raises(ValueError, match="reduce_volume"):
        DetectorConvergenceCondition("det", _wc_for_spp(10)).setup((0, _det_arrays(np.zeros((T,), F))), _CONFIG, None)
    with pytest.raises(ValueError, match="exactly the same"):
        DetectorConvergenceCondition("det", _wc_for_spp(10)).setup((0, _det_arrays(np.zeros((T - 1, 1), F))), _CONFIG, None)
    with pytest.raises(ValueError, match="prev_periods"):
        DetectorConvergenceCondition("det", _wc_for_spp(10), prev_periods=0).setup((0, ok), _CONFIG, None)
    with pytest.raises(ValueError, match="non-negative"):
        DetectorConvergenceCondition("det", _wc_for_spp(10), threshold=-1.0).setup((0, ok), _CONFIG, None)
    with pytest.raises(ValueError, match="min_steps must be larger"):
        DetectorConvergenceCondition("det", _wc_for_spp(10), prev_periods=4, min_steps=10).setup((0, ok), _CONFIG, None)
    with pytest.raises(ValueError, match="greater than the number of time steps"):
        DetectorConvergenceCondition("det", _wc_for_spp(T), prev_periods=4).setup((0, ok), _CONFIG, None)


def test_detector_convergence_truth_table():
    T = _CONFIG.time_steps_total
    spp, pp = 10, 4
    n = np.arange(T)
    periodic = np.sin(2 * np.pi * n / spp).astype(F)[:, None]          # converged signal
    growing = (np.sin(2 * np.pi * n / spp) * (1 + 0.05 * n)).astype(F)[:, None]
    for readings, expect_continue in ((periodic, False), (growing, True)):
        arrays = _det_arrays(readings)
        cond = DetectorConvergenceCondition("det", _wc_for_spp(spp), prev_periods=pp, threshold=1e-3).setup((0, arrays), _CONFIG, None)
        t = cond.min_steps + 7
        assert _ev(cond, cond.min_steps - 1, arrays) is True        # before min_steps: always continue
        assert _ev(cond, t, arrays) is expect_continue
        assert cond((t, arrays), _CONFIG, None) is expect_continue   # host mirror agrees with the oracle
        assert _ev(cond, T, arrays) is False                         # hard stop at time_steps_total
        a, b = cond.window(t, T)
        assert (a, b) == (t - (pp + 1) * spp, t - spp)


# ---- oracle runs (ref. tests/integration/fdtd/test_stop_conditions.py) --------------------------------
def test_oracle_energy_threshold_stops_a_decaying_pulse_early():
    objects, arrays, cfg = build_scene(shape=(10, 9, 12), thickness=3, source="pulse", detectors=("energy_reduce",), time=60e-15)
    T = cfg.time_steps_total
    full = yee.checkpointed_fdtd(arrays, objects, cfg)
    assert full[0] == T  # default TimeStepCondition: runs to the end
    trace = full[1].detector_states["energy_reduce"]["energy"][:, 0]
    e_end = float(np.sum(yee.compute_energy(full[1].fields.E, full[1].fields.H, full[1].inv_permittivities, full[1].inv_permeabilities)))
    thr = float(trace.max()) * 1e-2 * (e_end / float(trace[-1]))
    peak = int(np.argmax(trace))
    st = yee.checkpointed_fdtd(arrays, objects, cfg, stopping_condition=EnergyThresholdCondition(threshold=thr, min_steps=peak + 2))
    assert peak + 2 <= st[0] < T
    # stopped exactly when the energy first fell below the threshold: one more check would also stop
    cond = EnergyThresholdCondition(threshold=thr, min_steps=peak + 2).setup((0, arrays), cfg, objects)
    assert yee.evaluate_condition(cond, st, cfg, objects) is False
    # max_steps is a hard cut-off
    st2 = yee.checkpointed_fdtd(arrays, objects, cfg, stopping_condition=EnergyThresholdCondition(threshold=1e-300, min_steps=1, max_steps=17))
    assert st2[0] == 17


# ---- device side ------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(eps_tier=3, mu_tier=3, shape=(9, 7, 13))])
def test_device_energy_matches_oracle(kw):
    from fdtdx_b200.fdtd import get_plan
    from scenes import seed_fields

    objects, arrays, cfg = build_scene(**kw)
    seed_fields(arrays, seed=4)
    ref = float(np.sum(yee.compute_energy(arrays.fields.E, arrays.fields.H, arrays.inv_permittivities, arrays.inv_permeabilities).astype(np.float64)))
    dev = arrays.to_torch("cuda")
    got = float(get_plan(dev, objects, cfg).total_energy(dev).item())
    assert abs(got - ref) <= 1e-5 * abs(ref)


@pytest.mark.gpu
def test_energy_threshold_run_stops_where_the_oracle_stops():
    """A pulse in a CPML box decays; both sides must stop at the same step with the same fields."""
    objects, arrays, cfg = build_scene(shape=(14, 12, 16), thickness=4, source="pulse", detectors=("energy_reduce",), time=60e-15)
    T = cfg.time_steps_total
    probe = yee.checkpointed_fdtd(arrays, objects, cfg)
    e_end = float(np.sum(yee.compute_energy(probe[1].fields.E, probe[1].fields.H, probe[1].inv_permittivities, probe[1].inv_permeabilities)))
    trace = probe[1].detector_states["energy_reduce"]["energy"][:, 0]
    # the detector trace and sum(compute_energy) differ by a constant factor (cell volume); rescale
    thr = float(trace.max()) * 1e-2 * (e_end / float(trace[-1]))
    cond = EnergyThresholdCondition(threshold=thr, min_steps=int(np.argmax(trace)) + 2)
    st_o = yee.checkpointed_fdtd(arrays, objects, cfg, stopping_condition=cond)
    assert 0 < st_o[0] < T, (st_o[0], T, e_end, thr)
    st_g = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg, stopping_condition=cond)
    assert st_g[0] == st_o[0]
    assert rel_l2(st_g[1].fields.E.cpu().numpy(), st_o[1].fields.E) <= 1e-5
    assert rel_l2(st_g[1].fields.H.cpu().numpy(), st_o[1].fields.H) <= 1e-5


@pytest.mark.gpu
def test_detector_convergence_run_matches_oracle():
    objects, arrays, cfg = build_scene(shape=(14, 12, 16), thickness=4, source="plane_z", detectors=("energy_reduce",), time=40e-15)
    wc = next(iter(objects.sources)).wave_character
    half = fx.WaveCharacter(period=wc.get_period() / 2)
    cond = DetectorConvergenceCondition("energy_reduce", half, prev_periods=2, threshold=float(1e-22))
    st_o = yee.checkpointed_fdtd(arrays, objects, cfg, stopping_condition=cond)
    st_g = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg, stopping_condition=cond)
    assert st_g[0] == st_o[0]
    assert rel_l2(st_g[1].fields.E.cpu().numpy(), st_o[1].fields.E) <= 1e-5
