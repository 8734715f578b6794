"""Dispersive-material setup the Yee step consumes (host side, float64 NumPy).

The hot path only reads the ADE recurrence coefficient arrays ``c1..c4`` and, for TFSF sources that
sit in a dispersive medium, the filtered H-side temporal table (``tfsf.py:259-264``).  This module
derives both the way the reference does at setup time:

* ``pole_coefficients``      - ``fdtdx/dispersion.py:789-885`` (Lorentz / Drude / CCPR poles ->
  ``c1 = (2 - w0^2 dt^2)/D, c2 = -(1 - g dt/2)/D, c3 = (a dt^2 - b dt)/D, c4 = b dt/D``,
  ``D = 1 + g dt/2``),
* ``susceptibility``         - ``dispersion.py:988-1064`` (chi(w) recovered from the coefficients),
* ``effective_inv_permittivity`` - ``dispersion.py:1254-1307`` (carrier-frequency 1/Re(eps_inf+chi)),
* ``dispersive_H_filter``    - ``objects/sources/tfsf.py:21-140`` + ``dispersion.py:1067-1251``
  (s_H = irfft(rfft(s) * sqrt(eps(w)/eps(w_c))), DC gain 1, real Nyquist bin).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class LorentzPole:
    """chi(w) = delta_eps * w0^2 / (w0^2 - w^2 - i g w)  (``dispersion.py:372-414``)."""

    resonance_frequency: float
    damping: float
    delta_epsilon: float

    def abg(self):  # (a, b, gamma, omega_0): numerator a - i w b
        return self.delta_epsilon * self.resonance_frequency**2, 0.0, self.damping, self.resonance_frequency


@dataclass(frozen=True)
class DrudePole:
    """chi(w) = -wp^2 / (w^2 + i g w)  (``dispersion.py:417-456``)."""

    plasma_frequency: float
    damping: float

    def abg(self):
        return self.plasma_frequency**2, 0.0, self.damping, 0.0


def pole_coefficients(poles, dt: float):
    """(c1, c2, c3, c4) float64 arrays of shape (len(poles),)."""
    out = np.zeros((4, len(poles)))
    for i, p in enumerate(poles):
        a, b, g, w0 = p.abg()
        if (a != 0.0 or b != 0.0) and w0 * dt >= 2.0:
            raise ValueError(f"pole {i}: omega_0 * dt = {w0 * dt:.4g} >= 2 (ADE recurrence unstable)")
        D = 1.0 + 0.5 * g * dt
        out[:, i] = ((2.0 - (w0 * dt) ** 2) / D, -(1.0 - 0.5 * g * dt) / D, (a * dt * dt - b * dt) / D, b * dt / D)
    return out[0], out[1], out[2], out[3]


def coefficient_arrays(poles, dt: float, mask: np.ndarray):
    """Per-cell coefficient arrays (n_poles, 1, Nx, Ny, Nz) float32 for the cells selected by the
    boolean ``mask`` (zero elsewhere = no pole), as ``_init_arrays`` stores them."""
    c = pole_coefficients(poles, dt)
    arrs = []
    for k in range(4):
        a = np.zeros((len(poles), 1, *mask.shape), np.float32)
        for p in range(len(poles)):
            a[p, 0][mask] = np.float32(c[k][p])
        arrs.append(a)
    has_c4 = bool(np.any(c[3] != 0.0))
    return {"c1": arrs[0], "c2": arrs[1], "c3": arrs[2], "c4": arrs[3] if has_c4 else None}


def _pole_params(c1, c2, c3, c4):
    c1, c2, c3 = (np.asarray(x, np.float64) for x in (c1, c2, c3))
    c4 = np.zeros_like(c3) if c4 is None else np.asarray(c4, np.float64)
    live = (c1 != 0.0) | (c3 != 0.0) | (c4 != 0.0)
    den = np.where(1.0 - c2 == 0.0, 1.0, 1.0 - c2)
    g_dt = np.where(live, 2.0 * (1.0 + c2) / den, 0.0)
    D = 1.0 + 0.5 * g_dt
    w0sq = np.where(live, 2.0 - c1 * D, 0.0)
    a = np.where(live, (c3 + c4) * D, 0.0)
    b = np.where(live, c4 * D, 0.0)
    return live, g_dt, w0sq, a, b


def susceptibility(c1, c2, c3, omega, dt: float, c4=None):
    """chi(omega) summed over the pole axis; ``omega`` scalar or 1-D (prepended as axis 0)."""
    live, g_dt, w0sq, a, b = _pole_params(c1, c2, c3, c4)
    om = np.atleast_1d(np.asarray(omega, np.float64))
    wd = (om * dt).reshape((-1,) + (1,) * a.ndim)
    num = a[None] - 1j * wd * b[None]
    den = w0sq[None] - wd * wd - 1j * g_dt[None] * wd
    chi = np.where(live[None], num / np.where(live[None], den, 1.0 + 0.0j), 0.0 + 0.0j).sum(axis=1)
    return chi[0] if np.ndim(omega) == 0 else chi


def effective_inv_permittivity(inv_eps, c1, c2, c3, omega: float, dt: float, c4=None):
    """1 / Re(eps_inf + chi(omega)) per cell, dtype of ``inv_eps`` (diagonal tiers)."""
    inv_eps = np.asarray(inv_eps)
    chi = susceptibility(c1, c2, c3, omega, dt, c4)
    return (1.0 / (1.0 / inv_eps.astype(np.float64) + chi.real)).astype(inv_eps.dtype)


def dispersive_H_filter(raw_samples, dt: float, c1, c2, c3, inv_eps_inf, omega_c: float, c4=None) -> np.ndarray:
    """Broadband impedance-corrected H-side profile (one value per integer time step)."""
    raw = np.asarray(raw_samples, np.float64)
    n = raw.shape[0]
    c3n = np.asarray(c3)
    if c3n.size == 0 or (not np.any(c3n) and (c4 is None or not np.any(np.asarray(c4)))):
        return raw
    m = 1
    while m < 2 * n:
        m *= 2
    omegas = 2.0 * np.pi * np.fft.rfftfreq(m, d=dt)
    eps_inf = np.mean(1.0 / np.asarray(inv_eps_inf, np.float64), axis=0)

    def mean_eps(om):
        chi = susceptibility(c1, c2, c3, om, dt, c4)  # (M, C, *spatial)
        chi = chi.mean(axis=1)
        return (eps_inf[None] + chi).reshape(chi.shape[0], -1).mean(axis=1)

    eps_w = mean_eps(omegas)
    eps_c = complex(mean_eps(np.array([omega_c]))[0])
    padded = np.zeros(m)
    padded[:n] = raw
    G = np.sqrt(eps_w / eps_c)
    G[0] = 1.0
    G[-1] = G[-1].real
    return np.fft.irfft(np.fft.rfft(padded) * G, n=m)[:n]
