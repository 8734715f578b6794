"""Complex (lossy-mode) TFSF profiles: the imaginary part of the incident fields is injected in
quadrature with the carrier phase shifted by -pi/2 (objects/sources/tfsf.py:266-283, 366-383;
objects/sources/mode.py:212-222)."""

import numpy as np
import pytest

import fdtdx_b200 as fx
from oracle import yee
from scenes import build_scene, rel_l2


def _complexify(objects, seed=5):
    src = objects.sources[0]
    rng = np.random.default_rng(seed)
    src._E = (src._E + 1j * (0.4 * src._E + 0.05 * np.abs(src._E).max() * rng.standard_normal(src._E.shape))).astype(np.complex64)
    src._H = (src._H + 1j * (-0.3 * src._H + 0.05 * np.abs(src._H).max() * rng.standard_normal(src._H.shape))).astype(np.complex64)
    return src


def test_oracle_quadrature_of_a_phase_rotated_profile():
    """Known answer: a profile multiplied by exp(-i phi) and injected in quadrature equals the real
    profile injected with the carrier delayed by phi:  Re(F) cos(wt) + Im(F) cos(wt - pi/2) with
    F = F0 (cos phi - i sin phi)  ->  F0 cos(wt + phi)."""
    phi = 0.7
    kw = dict(source="plane_x", time=6e-15, detectors=())
    o1, a1, c1 = build_scene(**kw)
    s1 = o1.sources[0]
    s1._E = (s1._E * np.exp(-1j * phi)).astype(np.complex64)
    s1._H = (s1._H * np.exp(-1j * phi)).astype(np.complex64)
    o2, a2, c2 = build_scene(**kw)
    s2 = o2.sources[0]
    s2.wave_character = fx.WaveCharacter(wavelength=s2.wave_character.get_wavelength(), phase_shift=phi)
    # compare after the 4-period start-up ramp is irrelevant: both use the same ramp, only the carrier differs
    r1 = yee.checkpointed_fdtd(a1, o1, c1)[1]
    r2 = yee.checkpointed_fdtd(a2, o2, c2)[1]
    assert np.abs(r2.fields.E).max() > 0
    assert rel_l2(r1.fields.E, r2.fields.E) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("source,extra", [("plane_x", {}), ("plane_z", {}), ("pulse", {}), ("plane_x", {"eps_tier": 9}), ("plane_y", {"mu_tier": 3, "eps_tier": 3})])
def test_complex_mode_source_quadrature(source, extra):
    objects, arrays, cfg = build_scene(source=source, time=6e-15, detectors=("field",), shape=(14, 12, 16), **extra)
    _complexify(objects)
    ref = yee.checkpointed_fdtd(arrays, objects, cfg)[1]
    t_end, out = fx.run_fdtd(arrays.to_torch("cuda"), objects, cfg)
    assert np.abs(ref.fields.E).max() > 0
    assert rel_l2(out.fields.E.cpu().numpy(), ref.fields.E) <= 1e-5
    assert rel_l2(out.fields.H.cpu().numpy(), ref.fields.H) <= 1e-5
    # the imaginary part matters: dropping it changes the field by far more than the tolerance
    objects2, arrays2, cfg2 = build_scene(source=source, time=6e-15, detectors=("field",), shape=(14, 12, 16), **extra)
    real_only = yee.checkpointed_fdtd(arrays2, objects2, cfg2)[1]
    assert rel_l2(real_only.fields.E, ref.fields.E) > 1e-2
