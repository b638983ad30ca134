"""The analytic helpers of the reference (helpers.py:10-140) under their own names: laser tuning for a target
modulation depth, the Cramer-Rao bound on m, the SNR <-> noise-density conversion, the model Jacobian.  Host
arithmetic on a handful of numbers; Bessel values come from the package's own Miller recurrence (montecarlo.bessel_j)."""
from __future__ import annotations

import numpy as np

from .montecarlo import bessel_j, crlb_sigma_m
from .physics import SPEED_OF_LIGHT


def set_laser_df_for_effect(laser, ifo, m):
    """Set ``laser.df`` such that the interferometer sees modulation depth ``m`` (helpers.py:10-14)."""
    opd = np.abs(ifo.meas_arml - ifo.ref_arml)
    laser.df = (m * SPEED_OF_LIGHT) / (2 * np.pi * opd)


def calculate_crlb_for_m(m_true, ndata, snr_db, buffer_size):
    """Cramer-Rao lower bound on sigma(m) (helpers.py:16-45)."""
    return crlb_sigma_m(m_true, ndata, snr_db, buffer_size)


def snr_to_asd(snr_db, f_samp):
    """Amplitude-noise density equivalent to an SNR for a unit cosine (power 0.5), white noise over f_samp / 2
    (helpers.py:47-58)."""
    noise_power = 0.5 / 10 ** (snr_db / 10.0)
    return np.sqrt(noise_power / (f_samp / 2.0))


def calculate_jacobian(ndata, param):
    """2N x 4 Jacobian of the harmonic model in column order [a, m, phi, psi] (helpers.py:60-96; fit.py:123-144)."""
    a, m, phi, psi = param
    j = np.arange(1, ndata + 1)
    bes = bessel_j(ndata + 1, m)
    B, dB = bes[1:ndata + 1], 0.5 * (bes[0:ndata] - bes[2:ndata + 2])
    quarter = np.cos(phi + j * np.pi / 2.0)
    dquarter = np.cos(phi + j * np.pi / 2.0 + np.pi / 2.0)
    c, s = np.cos(j * psi), np.sin(j * psi)
    env = a * quarter * B
    jac = np.zeros((2 * ndata, 4))
    if a != 0:
        jac[:ndata, 0] = env * c / a
        jac[ndata:, 0] = -env * s / a
    jac[:ndata, 1] = a * quarter * dB * c
    jac[ndata:, 1] = -(a * quarter * dB) * s
    jac[:ndata, 2] = a * dquarter * B * c
    jac[ndata:, 2] = -(a * dquarter * B) * s
    jac[:ndata, 3] = env * -s * j
    jac[ndata:, 3] = -env * c * j
    return jac


def calculate_m_precision(m_range, ndata, snr_db):
    """Statistical uncertainty of m over ``m_range`` at phi = pi/4, unit amplitude, I/Q noise variance 1 / SNR_amplitude^2
    (helpers.py:98-140)."""
    noise_variance = (1.0 / 10 ** (snr_db / 20.0)) ** 2
    out = []
    for m_true in m_range:
        jac = calculate_jacobian(ndata, np.array([1.0, m_true, np.pi / 4, 0.0]))
        try:
            cov = noise_variance * np.linalg.inv(jac.T @ jac)
            out.append(np.sqrt(cov[1, 1]))
        except np.linalg.LinAlgError:
            out.append(np.inf)
    return np.array(out)
