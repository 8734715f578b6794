"""Physical constants, mirroring the reference's ``fdtdx/constants.py:13-23``."""

import math

c: float = 299792458.0
mu0: float = 4e-7 * math.pi
eps0: float = 1.0 / (mu0 * c**2)
eta0: float = mu0 * c


def wavelength_to_period(wavelength: float) -> float:
    return wavelength / c
