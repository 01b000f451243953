"""Lloyd-Max quantiser design and Bussgang gain (host side).

Function surface of the reference's ``modules/lloyd_max_quantizer.py`` (lines 6-89).  The design
iterates centroid / midpoint updates on the positive half line for a zero-mean Gaussian of
variance ``(1 + sigma^2)/2``; where the reference integrates numerically with ``scipy.integrate.quad``
this implementation uses the closed forms of the truncated-Gaussian moments (erf / exp), which agree
with the quadrature to ~1e-10, far inside the reference's own stopping tolerance (1e-5).
"""
import math

import numpy as np
from scipy.special import erf


def get_rho_lloyd(snr_dB, n_bits):
    """High-resolution distortion approximation (reference :6-7)."""
    return n_bits * 2 ** (-2 * n_bits)


def bussgang_diag(n_bits, var, quantizer):
    """Diagonal Bussgang gain for arbitrary labels: ``sum_j l_j (e^{-t_j^2/v} - e^{-t_{j+1}^2/v}) / sqrt(pi v)``
    with ``t_0 = -inf``, ``t_{2^b} = +inf`` (reference :10-21)."""
    var = np.asarray(var, dtype=float)
    thr = np.concatenate([[-np.inf], np.asarray(quantizer[0], dtype=float), [np.inf]])
    lab = np.asarray(quantizer[1], dtype=float)
    inv = 1 / var
    levels = int(2 ** n_bits)
    acc = -lab[0] * np.exp(-thr[1] ** 2 * inv)
    acc = acc + lab[levels - 1] * np.exp(-thr[levels - 1] ** 2 * inv)
    for j in range(1, levels - 1):
        acc = acc + lab[j] * (np.exp(-thr[j] ** 2 * inv) - np.exp(-thr[j + 1] ** 2 * inv))
    return acc / (math.sqrt(math.pi) * np.sqrt(var))


def get_Bussgang_matrix(n_bits, Cy, quantizer):
    return np.diag(bussgang_diag(n_bits, np.real(np.diag(np.asarray(Cy))), quantizer).astype(complex))


def _gauss_partial_moments(a, b, sd):
    """``(int_a^b p, int_a^b x p, int_a^b x^2 p)`` for ``p = N(0, sd^2)``; ``b`` may be +inf."""
    s2 = sd * math.sqrt(2.0)
    Pa, Pb = 0.5 * (1 + erf(a / s2)), (1.0 if np.isinf(b) else 0.5 * (1 + erf(b / s2)))
    pdf = lambda x: 0.0 if np.isinf(x) else math.exp(-0.5 * (x / sd) ** 2) / (sd * math.sqrt(2 * math.pi))
    fa, fb = pdf(a), pdf(b)
    m0 = Pb - Pa
    m1 = sd ** 2 * (fa - fb)
    m2 = sd ** 2 * m0 + sd ** 2 * (a * fa - (0.0 if np.isinf(b) else b * fb))
    return m0, m1, m2


def lloyd_max_quantizer(levels, mean, variance, max_iter=200, pk_gmm=None):
    """Lloyd-Max design for the positive half of a zero-mean Gaussian (reference :40-89).

    Returns ``(intervals[levels+1], centroids[levels], rho)`` like the reference: ``intervals[0] = 0``,
    ``intervals[-1] = inf``; ``rho`` is the distortion accumulated over the positive half line.
    """
    if pk_gmm is not None or mean != 0:
        raise NotImplementedError('only the single zero-mean Gaussian design is implemented')
    variance = float(np.real(variance))
    sd = math.sqrt(variance)
    upper = float(np.clip(3 * variance, 0, 100))
    edges = np.append(np.linspace(0.0, upper, levels), np.inf)
    cent = np.zeros(levels)
    for _ in range(max_iter):
        before = edges[:-1].copy()
        for j in range(levels):
            m0, m1, _ = _gauss_partial_moments(edges[j], edges[j + 1], sd)
            cent[j] = m1 / m0 if m0 > 0 else 0.5 * (edges[j] + edges[j + 1])
        edges[1:-1] = 0.5 * (cent[1:] + cent[:-1])
        if np.linalg.norm(before - edges[:-1]) < 1e-5:
            break
    rho = 0.0
    for j in range(levels):
        m0, m1, m2 = _gauss_partial_moments(edges[j], edges[j + 1], sd)
        rho += m2 - 2 * cent[j] * m1 + cent[j] ** 2 * m0
    return edges, cent, rho


def load_quantizer(snr, n_bits, sigmas_gmm=None, pk_gmm=None):
    """Symmetric Lloyd-Max table for per-dimension input variance ``(1 + sigma^2)/2`` (reference :24-37)."""
    if sigmas_gmm is not None:
        raise NotImplementedError('GMM-input Lloyd-Max design is outside the inference path')
    edges, cent, rho = lloyd_max_quantizer(int(2 ** n_bits / 2), 0, 0.5 * (1 + 10 ** (-snr / 10)))
    pos = edges[:-1]
    thresholds = np.concatenate([-pos[:0:-1], pos])
    labels = np.concatenate([-cent[::-1], cent])
    return {snr: (thresholds, labels, rho)}
