"""Uniform mid-rise quantiser statistics (host side).

Mirrors the function surface of the reference's ``modules/uniform_quantizer.py`` (numpy parts,
lines 6-72, 114-128, 149-173): Max's step table, distortion factors, the SNR-scaled step, the
diagonal Bussgang gain, the quantised variance and the arcsine / mean-gain ``C_r``.  These run
once per (SNR, bit width) on the host in float64; the per-sample work is in the CUDA kernels.
"""
import math
import warnings

import numpy as np

# J. Max, "Quantizing for minimum distortion", table 2 (N(0,1) input): step and distortion per bit width
_MAX_STEP = (None, 1.596, 0.9957, 0.5860, 0.3352, 0.1881, 0.1041, 0.0569, 0.0308)
_MAX_RHO = (None, 1 - 2 / np.pi, 0.11885, 0.037440, 0.011535, 0.0034914, 0.00104, 0.00030433, 0.00008769)


def standard_quantization_step(n_bits):
    """Optimal uniform step for N(0,1) input (reference uniform_quantizer.py:6-23)."""
    if n_bits <= 8:
        return _MAX_STEP[n_bits]
    warnings.warn('Optimal standard step size is unknown and thus approximated!')
    return 4 * np.sqrt(n_bits) * 2 ** (-n_bits)          # Hui & Neuhoff asymptote


def get_uniform_quant_step(snr_dB, n_bits):
    """Step scaled to the per-real-dimension input power (1 + sigma^2)/2 (reference :44-45)."""
    return np.sqrt((1 + 10 ** (-snr_dB / 10)) / 2) * standard_quantization_step(n_bits)


def get_rho_uniform(snr_dB, n_bits):
    """Granular + overload distortion approximation (reference :52-57)."""
    step = get_uniform_quant_step(snr_dB, n_bits)
    rho = step ** 2 / 12
    rho += np.exp(-2 ** (2 * n_bits - 3) * step ** 2) / (2 ** (n_bits - 1.5) * step) ** 3 / np.sqrt(np.pi)
    return rho


def standard_distortion_fac(n_bits):
    """Distortion factor of the optimal uniform quantiser (reference :26-41)."""
    if n_bits <= 8:
        return _MAX_RHO[n_bits]
    warnings.warn('Optimal standard distortion factor is unknown and thus approximated!')
    return get_rho_uniform(np.inf, n_bits)


def bussgang_diag(snr_dB, n_bits, var):
    """Diagonal of the Bussgang gain for per-antenna variances ``var`` (real array, any shape).

    1 bit: ``sqrt(2/pi)/sqrt(var)``; b bits: ``step/sqrt(pi var) * sum_j exp(-step^2 (j - 2^(b-1))^2 / var)``
    for ``j = 1 .. 2^b - 1`` (reference :60-72).
    """
    var = np.asarray(var, dtype=float)
    if n_bits == np.inf:
        return np.ones_like(var)
    if n_bits == 1:
        return math.sqrt(2 / math.pi) * (1 / np.sqrt(var))
    step = get_uniform_quant_step(snr_dB, n_bits)
    inv = 1 / var
    levels = int(2 ** n_bits)
    acc = np.zeros_like(var)
    for j in range(1, levels):
        acc = acc + np.exp(-step ** 2 * (j - levels / 2) ** 2 * inv)
    return acc * (step / math.sqrt(math.pi) / np.sqrt(var))


def get_Bussgang_matrix(snr_dB, n_bits, Cy):
    """Diagonal Bussgang gain matrix for the covariance ``Cy`` (reference :60-72)."""
    Cy = np.asarray(Cy)
    if n_bits == np.inf:
        return np.eye(Cy.shape[-1])
    return np.diag(bussgang_diag(snr_dB, n_bits, np.real(np.diag(Cy))).astype(complex))


def get_quantized_variance(sigma2, quantizer):
    """Variance of the quantiser output for complex input variance ``sigma2`` (reference :114-128)."""
    from scipy.stats import norm
    per_dim = np.atleast_1d(np.asarray(sigma2, dtype=complex).real / 2)
    sd = np.sqrt(per_dim)
    thr, lab = np.asarray(quantizer[0]), np.asarray(quantizer[1])
    cdf = norm.cdf(thr[:, None] / sd[None, :])                       # [n_thr, n]
    edges = np.concatenate([np.zeros((1, sd.size)), cdf, np.ones((1, sd.size))], axis=0)
    res = np.sum(lab[:, None] ** 2 * np.diff(edges, axis=0), axis=0)
    return 2 * np.squeeze(res)


def arcsine_law(Cy):
    """1-bit ``C_r = 2/pi (asin(Re rho) + j asin(Im rho))`` with ``rho`` the normalised ``Cy``."""
    Cy = np.asarray(Cy)
    s = 1 / np.sqrt(np.real(np.diagonal(Cy, axis1=-2, axis2=-1)))
    rho = s[..., :, None] * Cy * s[..., None, :]
    return 2 / np.pi * (np.arcsin(np.clip(rho.real, -1.0, 1.0)) + 1j * np.arcsin(np.clip(rho.imag, -1.0, 1.0)))


def get_Cr(Cy, n_bits, snr=None, quantizer=None):
    """Covariance of the quantised observation used by the scripts' rate bounds (reference :149-173)."""
    Cy = np.asarray(Cy)
    if n_bits == 1:
        return arcsine_law(Cy)
    if n_bits == np.inf:
        return Cy
    if Cy.ndim != 2:
        raise ValueError('get_Cr: one covariance matrix at a time for n_bits > 1')
    gain = bussgang_diag(snr, n_bits, np.real(np.diag(Cy)))
    Cr = np.array(np.mean(gain) ** 2 * Cy, dtype=Cy.dtype)
    np.fill_diagonal(Cr, get_quantized_variance(np.diag(Cy), quantizer))
    return Cr
