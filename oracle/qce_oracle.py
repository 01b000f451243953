"""CPU oracle for the Bussgang-GMM / Bussgang-MFA inference path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``quantized_channel_estimation_b200/`` imports it.

It is a numpy (complex128) restatement of the reference algorithm
(benediktfesl/Quantized_Channel_Estimation); every function cites the reference
``file:line`` it follows (paths relative to the reference root).  Where the
reference runs a Python ``(sample, component)`` double loop that recomputes an
``N x N x N`` product per iteration (modules/gmm_cplx_bussgang.py:223-228) the oracle
evaluates the same expression batched over samples -- identical maths, different
summation order (differences are O(1e-14) relative).

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against *outputs of the reference itself*, produced by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` unmodified) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every
oracle function against those fixtures.

Third-party arithmetic used exactly like the reference does: numpy/LAPACK
(``matmul``, ``linalg.pinv``, ``linalg.slogdet``, ``digitize``, ``sign``, ``arcsin``),
``scipy.linalg.cholesky/solve_triangular/pinvh``, ``scipy.special.logsumexp``,
``scipy.integrate.quad`` + ``scipy.stats.norm`` (Lloyd-Max design only).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
from scipy import integrate
from scipy.special import logsumexp
from scipy.stats import norm

# --------------------------------------------------------------------------------------
# Quantiser tables and Bussgang statistics
# --------------------------------------------------------------------------------------

#: Max's optimal uniform step sizes for N(0,1) input -- modules/uniform_quantizer.py:11-12
MAX_UNIFORM_STEP = {1: 1.596, 2: 0.9957, 3: 0.5860, 4: 0.3352, 5: 0.1881, 6: 0.1041,
                    7: 0.0569, 8: 0.0308}
#: distortion factors -- modules/uniform_quantizer.py:31-32
MAX_UNIFORM_RHO = {1: 1 - 2 / np.pi, 2: 0.11885, 3: 0.037440, 4: 0.011535, 5: 0.0034914,
                   6: 0.00104, 7: 0.00030433, 8: 0.00008769}


def standard_quantization_step(n_bits):
    """modules/uniform_quantizer.py:6-23."""
    if n_bits <= 8:
        return MAX_UNIFORM_STEP[n_bits]
    return 4 * np.sqrt(n_bits) * 2 ** (-n_bits)


def get_uniform_quant_step(snr_dB, n_bits):
    """modules/uniform_quantizer.py:44-45."""
    return np.sqrt((1 + 10 ** (-snr_dB / 10)) / 2) * standard_quantization_step(n_bits)


def get_rho_uniform(snr_dB, n_bits):
    """modules/uniform_quantizer.py:52-57."""
    delt = get_uniform_quant_step(snr_dB, n_bits)
    rho = delt ** 2 / 12
    rho += np.exp(-2 ** (2 * n_bits - 3) * delt ** 2) / (2 ** (n_bits - 1.5) * delt) ** 3 / np.sqrt(np.pi)
    return rho


def standard_distortion_fac(n_bits):
    """modules/uniform_quantizer.py:26-41."""
    if n_bits <= 8:
        return MAX_UNIFORM_RHO[n_bits]
    return get_rho_uniform(np.inf, n_bits)


def uniform_bussgang_diag(snr_dB, n_bits, cy_diag):
    """Diagonal of ``uniform_quantizer.get_Bussgang_matrix`` (modules/uniform_quantizer.py:60-72).

    ``cy_diag`` is the (complex, zero-imaginary) diagonal of ``C_y``; returns a complex vector.
    """
    cy_diag = np.asarray(cy_diag)
    if n_bits == np.inf:
        return np.ones(cy_diag.shape[-1])
    if n_bits == 1:
        return np.sqrt(2 / np.pi) * 1 / np.sqrt(cy_diag)
    delta = get_uniform_quant_step(snr_dB, n_bits)
    inv = 1 / cy_diag
    b = np.zeros(cy_diag.shape, dtype=complex)
    for i in range(1, int(2 ** n_bits)):
        b += np.exp(-delta ** 2 * (i - 2 ** n_bits / 2) ** 2 * inv)
    b *= delta / np.sqrt(np.pi) / np.sqrt(cy_diag)
    return b


def lloyd_bussgang_diag(n_bits, cy_diag, quantizer):
    """Diagonal of ``lloyd_max_quantizer.get_Bussgang_matrix`` (modules/lloyd_max_quantizer.py:10-21)."""
    tau = list(quantizer[0])
    labels = list(quantizer[1])
    tau.insert(0, -np.inf)
    tau.append(np.inf)
    inv = 1 / np.asarray(cy_diag)
    b = -labels[0] * np.exp(-tau[1] ** 2 * inv)
    b = b + labels[int(2 ** n_bits - 1)] * np.exp(-tau[int(2 ** n_bits - 1)] ** 2 * inv)
    for i in range(1, int(2 ** n_bits - 1)):
        b = b + labels[i] * (np.exp(-tau[i] ** 2 * inv) - np.exp(-tau[i + 1] ** 2 * inv))
    b = b / (np.sqrt(np.pi) * np.sqrt(np.asarray(cy_diag)))
    return b


def get_rho_lloyd(snr_dB, n_bits):
    """modules/lloyd_max_quantizer.py:6-7."""
    return n_bits * 2 ** (-2 * n_bits)


def lloyd_max_quantizer(levels, mean, variance, max_iter=200):
    """Lloyd-Max design on the positive half line (modules/lloyd_max_quantizer.py:40-89,
    single-Gaussian branch ``pk_gmm is None``)."""
    max_int = np.clip(3 * np.max(variance), 0, 100)
    intervals = np.zeros([levels + 1])
    intervals[:-1] = np.linspace(0., max_int, levels)
    intervals[-1] = np.inf
    centroids = np.zeros(levels)
    thresh = 1e-5
    sd = variance ** 0.5
    for _ in range(max_iter):
        prev = np.copy(intervals)
        for j in range(levels):
            try:
                num = integrate.quad(lambda x: x * norm.pdf(x, mean, sd), intervals[j], intervals[j + 1])[0]
                den = integrate.quad(lambda x: norm.pdf(x, mean, sd), intervals[j], intervals[j + 1])[0]
                centroids[j] = num / den
            except ZeroDivisionError:
                centroids[j] = (intervals[j] + intervals[j + 1]) / 2
        for j in range(levels - 1):
            intervals[j + 1] = (centroids[j + 1] + centroids[j]) / 2.
        if np.linalg.norm(prev[:-1] - intervals[:-1]) < thresh:
            break
    rho = 0.0
    for j in range(levels):
        rho += integrate.quad(lambda x: (x - centroids[j]) ** 2 * norm.pdf(x, mean, sd),
                              intervals[j], intervals[j + 1])[0]
    return intervals, centroids, rho


def load_quantizer(snr, n_bits):
    """modules/lloyd_max_quantizer.py:24-37 (``sigmas_gmm is None`` branch)."""
    sigma2 = 10 ** (-snr / 10)
    input_var = 0.5 * (1 + sigma2)
    thresholds, labels, rho = lloyd_max_quantizer(levels=int((2 ** n_bits) / 2), mean=0,
                                                  variance=np.real(input_var))
    thresholds = thresholds[:-1]
    thresholds = np.concatenate((np.flip(-thresholds[1:]), thresholds), axis=0)
    labels = np.concatenate((np.flip(-labels), labels), axis=0)
    return {snr: (thresholds, labels, rho)}


def get_quantizer(snrs, n_bits, quantizer_type='uniform'):
    """Threshold / label tables per SNR -- modules/utils.py:531-562 (and the serial twin
    ``get_quantizer_gauss`` :565-590; the reference's Pool is only a scheduling detail)."""
    quantizer = dict()
    if n_bits == 'inf' or n_bits == np.inf or n_bits == 1:
        for snr in snrs:
            quantizer[snr] = (None, None, None)
        return quantizer
    if quantizer_type == 'uniform':
        for snr in snrs:
            delta = get_uniform_quant_step(snr, n_bits)
            thresholds = np.zeros([int(2 ** n_bits - 1)])
            for nb in range(int((2 ** n_bits - 2) / 2)):
                thresholds[nb] = -((2 ** n_bits - 2) / 2 - nb) * delta
                thresholds[-nb - 1] = ((2 ** n_bits - 2) / 2 - nb) * delta
            labels = np.zeros([int(2 ** n_bits)])
            for nb in range(int(2 ** n_bits - 1)):
                labels[nb] = thresholds[nb] - delta / 2
            labels[-1] = thresholds[-1] + delta / 2
            quantizer[snr] = (thresholds, labels, None)
    elif quantizer_type == 'lloyd':
        for snr in snrs:
            quantizer[snr] = load_quantizer(snr, n_bits)[snr]
    else:
        raise NotImplementedError(f'Quantizer type {quantizer_type} not implemented!')
    return quantizer


def get_quantized_variance(sigma2, quantizer):
    """modules/uniform_quantizer.py:114-128."""
    sigma2 = np.asarray(sigma2) / 2
    thresh, labels = quantizer[0], quantizer[1]
    if sigma2.ndim < 2:
        sigma2 = np.expand_dims(sigma2, 0)
    sd = np.sqrt(sigma2)
    res = labels[0] ** 2 * norm.cdf(thresh[0] / sd)
    res = res + labels[-1] ** 2 * (1 - norm.cdf(thresh[-1] / sd))
    for i in range(1, labels.shape[0] - 1):
        res = res + labels[i] ** 2 * (norm.cdf(thresh[i] / sd) - norm.cdf(thresh[i - 1] / sd))
    return 2 * np.squeeze(res)


def uniform_get_bussgang_matrix(snr_dB, n_bits, Cy):
    """``uniform_quantizer.get_Bussgang_matrix`` as the full diagonal matrix (:60-72)."""
    return np.diag(uniform_bussgang_diag(snr_dB, n_bits, np.diag(Cy)).astype(complex))


def lloyd_get_bussgang_matrix(n_bits, Cy, quantizer):
    """``lloyd_max_quantizer.get_Bussgang_matrix`` as the full diagonal matrix (:10-21)."""
    return np.diag(lloyd_bussgang_diag(n_bits, np.diag(Cy), quantizer))


def get_Cr(Cy, n_bits, snr=None, quantizer=None):
    """modules/uniform_quantizer.py:149-173 for a single covariance ``Cy [N,N]`` (the scripts'
    rate-bound helper: arcsine law for 1 bit, mean-gain scaling with the exact quantised
    variance on the diagonal otherwise)."""
    if n_bits == 1:
        return _quantised_cov(Cy, None, 1)
    if n_bits == np.inf:
        return Cy
    b = uniform_bussgang_diag(snr, n_bits, np.diag(Cy))
    Cr = (np.mean(b) ** 2 * Cy).astype(Cy.dtype)
    np.fill_diagonal(Cr, get_quantized_variance(np.diag(Cy), quantizer))
    return Cr


# --------------------------------------------------------------------------------------
# Quantiser application / observation synthesis
# --------------------------------------------------------------------------------------

def crandn(*shape, rng):
    """modules/utils.py:13-14 (with an explicit generator instead of the module global)."""
    return np.sqrt(0.5) * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def quant(inp, n_bits=1, thresholds=None, quant_labels=None):
    """modules/utils.py:189-203.  1-bit: ``1/sqrt(2) (sign Re + j sign Im)``; b-bit:
    ``np.digitize`` (right=False) against the thresholds, then label lookup."""
    if n_bits == 1:
        return 1 / np.sqrt(2) * (np.sign(np.real(inp)) + 1j * np.sign(np.imag(inp)))
    idx_r = np.digitize(np.real(inp), thresholds)
    idx_i = np.digitize(np.imag(inp), thresholds)
    return quant_labels[idx_r] + 1j * quant_labels[idx_i]


def quant_codes(inp, n_bits=1, thresholds=None):
    """Integer level index per real dimension (what ``np.digitize`` returns at
    modules/utils.py:195-196; for 1 bit: 0 for negative, 1 for zero, 2 for positive, 3 for NaN,
    so that ``np.sign`` is ``code - 1``).  Returned as ``uint8 [..., 2]`` (re, im)."""
    re, im = np.real(inp), np.imag(inp)
    if n_bits == 1:
        with np.errstate(invalid='ignore'):
            cr = np.where(np.isnan(re), 3, np.sign(re) + 1).astype(np.uint8)   # NaN -> 3 (np.sign(nan) = nan)
            ci = np.where(np.isnan(im), 3, np.sign(im) + 1).astype(np.uint8)
    else:
        cr = np.digitize(re, thresholds).astype(np.uint8)
        ci = np.digitize(im, thresholds).astype(np.uint8)
    return np.stack([cr, ci], axis=-1)


def observe(h, snr_dB, noise, A=None):
    """Unquantised pilot ``y = A h + 10^(-snr/20) n`` -- modules/utils.py:241-247 with the
    noise draw ``n = crandn(...)`` supplied by the caller (so both implementations see the
    same draw).  Two roundings: ``s*n`` then ``+``, no fused multiply-add."""
    h = np.asarray(h)
    if A is None:
        y = h.astype(complex, copy=True) if not np.iscomplexobj(h) else h.astype(np.result_type(h.dtype, np.complex64), copy=True)
        y = y.astype(np.complex128)
    else:
        y = np.squeeze(np.matmul(A, np.expand_dims(h, 2)), axis=2).astype(np.complex128)
    y = y + 10 ** (-snr_dB / 20) * noise
    return y


def get_observation_nbit(h, snr, noise, A=None, n_bits=1, thresholds=None, labels=None):
    """modules/utils.py:241-251."""
    y = observe(h, snr, noise, A)
    if n_bits == 'inf' or n_bits == np.inf:
        return y
    return quant(y, n_bits, thresholds, labels)


# --------------------------------------------------------------------------------------
# Per-component precompute (Bussgang gain, C_r, whitening factor, inverse)
# --------------------------------------------------------------------------------------

def _bussgang_diag(cy_diag, snr_dB, n_bits, quantizer_type, quantizer):
    """The ``A_buss`` selection of modules/gmm_cplx_bussgang.py:274-284 / mofa:171-180."""
    if n_bits == 1:
        return np.sqrt(2 / np.pi) * (1 / np.sqrt(cy_diag))
    if n_bits == np.inf:
        return np.ones_like(cy_diag)
    if quantizer_type == 'uniform':
        return uniform_bussgang_diag(snr_dB, n_bits, cy_diag)
    if quantizer_type == 'lloyd':
        return lloyd_bussgang_diag(n_bits, cy_diag, quantizer)
    raise NotImplementedError(quantizer_type)


def _quantised_cov(cy, b_diag, n_bits):
    """``C_r`` from ``C_y`` -- arcsine law (gmm:292-301; mofa:189-198), beta-model
    (gmm:304-307; mofa:199-202) or identity for infinite resolution (gmm:302-303)."""
    if n_bits == 1:
        psi = np.real(np.diag(1 / np.sqrt(np.diag(cy))))
        inner_real = np.real(psi @ np.real(cy) @ psi)
        inner_imag = np.real(psi @ np.imag(cy) @ psi)
        inner_real = np.clip(inner_real, -1.0, 1.0)
        inner_imag = np.clip(inner_imag, -1.0, 1.0)
        return 2 / np.pi * (np.arcsin(inner_real) + 1j * np.arcsin(inner_imag))
    if n_bits == np.inf:
        return cy.copy()
    beta = np.clip(np.real(np.mean(b_diag)), 0, 1)
    return beta ** 2 * cy + (1 - beta ** 2) * np.diag(np.diag(cy))


def compute_precision_cholesky(covariances):
    """'full' branch of modules/gmm_cplx_bussgang.py:15-47: ``P_k = (L_k^-1)^H``."""
    msg = ("Fitting the mixture model failed because some components have "
           "ill-defined empirical covariance (for instance caused by singleton "
           "or collapsed samples). Try to decrease the number of components, "
           "or increase reg_covar.")
    K, N, _ = covariances.shape
    out = np.empty((K, N, N), dtype=complex)
    for k in range(K):
        try:
            chol = scipy.linalg.cholesky(covariances[k], lower=True)
        except scipy.linalg.LinAlgError:
            raise ValueError(msg)
        out[k] = scipy.linalg.solve_triangular(chol, np.eye(N), lower=True).T.conj()
    return out


def gmm_prepare(means, covs, A, snr_dB, n_bits=1, quantizer_type='uniform', quantizer=None):
    """``Gmm_nbit._prepare_for_prediction`` -- modules/gmm_cplx_bussgang.py:246-328.

    Returns a dict with ``m_r [K,No]``, ``C_y``, ``C_r``, ``C_r_inv`` (np.linalg.pinv),
    ``prec_chol`` (``P_k``), ``b`` (Bussgang diagonals ``[K,No]``) and ``A_eff [K,No,N]``.
    """
    means = np.asarray(means, dtype=complex)
    covs = np.asarray(covs, dtype=complex)
    K = means.shape[0]
    sigma2 = 10 ** (-snr_dB / 10)
    Am = np.matmul(A, means[:, :, None])[:, :, 0]                     # gmm:256
    cy = np.matmul(np.matmul(A, covs), A.conj().T)                    # gmm:268
    cy = cy + sigma2 * np.eye(cy.shape[-1])                           # gmm:269-271
    b = np.stack([_bussgang_diag(np.diag(cy[k]), snr_dB, n_bits, quantizer_type, quantizer)
                  for k in range(K)]).astype(complex)                 # gmm:274-284
    m_r = b * Am                                                      # gmm:287-288
    cr = np.stack([_quantised_cov(cy[k], b[k], n_bits) for k in range(K)])   # gmm:291-307
    prec_chol = compute_precision_cholesky(cr)                        # gmm:310
    cr_inv = np.stack([np.linalg.pinv(cr[k]) for k in range(K)])      # gmm:321-323
    a_eff = b[:, :, None] * A[None, :, :]                             # gmm:326  (diag(b) @ A)
    return dict(m_r=m_r, C_y=cy, C_r=cr, C_r_inv=cr_inv, prec_chol=prec_chol, b=b, A_eff=a_eff)


def gmm_weighted_log_prob(y, prep, weights):
    """``_estimate_weighted_log_prob`` -- modules/gmm_cplx_bussgang.py:369-435 ('full')."""
    B, No = y.shape
    P = prep['prec_chol']
    K = P.shape[0]
    log_det = np.real(np.sum(np.log(P.reshape(K, -1)[:, ::No + 1]), 1))      # gmm:55-82, :411
    quad = np.empty((B, K))
    for k in range(K):
        z = np.dot(y, P[k].conj()) - np.dot(prep['m_r'][k], P[k].conj())     # gmm:416
        quad[:, k] = np.sum(np.abs(z) ** 2, axis=1)                           # gmm:417
    return -(No * np.log(np.pi) + quad) + 2 * log_det + np.log(weights)      # gmm:435, :380-383


def gmm_predict_proba(y, prep, weights):
    """``predict_proba_cplx`` -- gmm:351-367, :632-656."""
    wlp = gmm_weighted_log_prob(y, prep, weights)
    return np.exp(wlp - logsumexp(wlp, axis=1)[:, None])


def _component_lmmse(y, means, covs, prep):
    """All per-component LMMSE estimates ``[K,B,N]`` -- the expression of
    modules/gmm_cplx_bussgang.py:331-332 with the arguments built at :225-228."""
    K = means.shape[0]
    out = np.empty((K, y.shape[0], means.shape[1]), dtype=complex)
    for k in range(K):
        a_eff = prep['A_eff'][k]
        c_hy = covs[k] @ a_eff.conj().T
        t = (y - a_eff @ means[k]) @ prep['C_r_inv'][k].T
        out[k] = means[k] + t @ c_hy.T
    return out


def _combine(proba, hk, mode, labels_fn):
    """Combination modes -- modules/gmm_cplx_bussgang.py:197-242 / mofa:125-158.

    ``proba [B,K]``, ``hk [K,B,N]``.  The reference tests ``isinstance(mode, int)`` first.
    """
    K, B, N = hk.shape
    h_est = np.zeros((B, N), dtype=complex)
    if isinstance(mode, int):
        if mode == 1:
            labels = labels_fn()
            return hk[labels, np.arange(B)]
        for b in range(B):
            idx = np.argsort(proba[b])[::-1][:mode]
            h_est[b] = (proba[b, idx][:, None] * hk[idx, b]).sum(0) / np.sum(proba[b, idx])
        return h_est
    if isinstance(mode, str) and mode == 'all':
        return np.einsum('bk,kbn->bn', proba, hk)
    for b in range(B):
        idx = np.argsort(proba[b])[::-1]
        nr = np.searchsorted(np.cumsum(proba[b, idx]), mode) + 1
        idx = idx[:nr]
        h_est[b] = (proba[b, idx][:, None] * hk[idx, b]).sum(0) / np.sum(proba[b, idx])
    return h_est


def gmm_estimate_from_y(means, covs, weights, y, snr_dB, A=None, n_summands_or_proba=1, n_bits=1,
                        quantizer_type='uniform', quantizer=None, return_aux=False):
    """``Gmm_nbit.estimate_from_y`` -- modules/gmm_cplx_bussgang.py:166-243."""
    means = np.asarray(means, dtype=complex)
    covs = np.asarray(covs, dtype=complex)
    y = np.asarray(y)
    if A is None:
        A = np.eye(means.shape[1], dtype=complex)
    prep = gmm_prepare(means, covs, A, snr_dB, n_bits, quantizer_type, quantizer)
    wlp = gmm_weighted_log_prob(y, prep, weights)
    proba = np.exp(wlp - logsumexp(wlp, axis=1)[:, None])
    hk = _component_lmmse(y, means, covs, prep)
    h_est = _combine(proba, hk, n_summands_or_proba, lambda: wlp.argmax(axis=1))   # gmm:349
    if return_aux:
        return h_est, dict(proba=proba, wlp=wlp, prep=prep)
    return h_est


# --------------------------------------------------------------------------------------
# MFA
# --------------------------------------------------------------------------------------

def mofa_prepare(means, covs, A, snr_dB, n_bits, quantizer_type='uniform', quantizer=None):
    """``Mofa._prepare_for_prediction`` -- modules/mofa_cplx_bussgang.py:162-212."""
    means = np.asarray(means, dtype=complex)
    covs = np.asarray(covs, dtype=complex)
    K = means.shape[0]
    sigma2 = 10 ** (-snr_dB / 10)
    m_y = np.matmul(A, means[:, :, None])[:, :, 0]                    # mofa:166
    cy = A @ covs @ A.conj().T + sigma2 * np.eye(A.shape[0])          # mofa:167-169
    b = np.stack([_bussgang_diag(np.diag(cy[k]), snr_dB, n_bits, quantizer_type, quantizer)
                  for k in range(K)]).astype(complex)                 # mofa:171-180
    m_r = b * m_y                                                     # mofa:183-184
    if n_bits != np.inf:
        cr = np.stack([_quantised_cov(cy[k], b[k], n_bits) for k in range(K)])   # mofa:187-204
    else:
        cr = cy
    cr_inv = np.stack([scipy.linalg.pinvh(cr[k]) for k in range(K)])  # mofa:205-207
    a_eff = b[:, :, None] * A[None, :, :]                             # mofa:210
    return dict(m_r=m_r, C_y=cy, C_r=cr, C_r_inv=cr_inv, b=b, A_eff=a_eff)


def mofa_log_resp_unnorm(y, prep, amps):
    """``log amps_k + _log_multi_gauss(k, y)`` -- mofa:346-348, :370-381.  Returns ``[K,B]``."""
    K = prep['C_r'].shape[0]
    out = np.zeros((K, y.shape[0]))
    for k in range(K):
        _, logdet = np.linalg.slogdet(prep['C_r'][k])
        x1 = (y - prep['m_r'][k]).T
        x2 = prep['C_r_inv'][k] @ x1
        p = np.sum(x1.conj() * x2, axis=0)
        out[k] = np.log(amps[k]) + np.real(-np.log(np.pi) * y.shape[1] - logdet - p)
    return out


def _log_sum(loglikes):
    """mofa:394-400."""
    a = np.max(loglikes, axis=0)
    return a + np.log(np.sum(np.exp(loglikes - a[None, :]), axis=0))


def mofa_predict_proba(y, prep, amps):
    """mofa:342-356."""
    logrs = mofa_log_resp_unnorm(y, prep, amps)
    return np.exp(logrs - _log_sum(logrs)[None, :]).T


def mofa_predict_proba_max(y, prep, amps):
    """mofa:359-366 -- argmax over ``exp`` of the *un-normalised* log-probabilities."""
    return np.exp(mofa_log_resp_unnorm(y, prep, amps)).argmax(axis=0)


def mofa_covs(lambdas, psis):
    """``C_k = Lambda_k Lambda_k^H + diag(psi_k)`` -- mofa:313-319."""
    covs = lambdas @ np.transpose(lambdas.conj(), [0, 2, 1])
    covs = covs + np.stack([np.diag(p) for p in psis])
    return covs


def mofa_estimate_from_y(means, covs, amps, y, snr_dB, A=None, n_summands_or_proba=1, n_bits=1,
                         quantizer_type='uniform', quantizer=None, return_aux=False):
    """``Mofa.estimate_from_y`` -- modules/mofa_cplx_bussgang.py:117-159, ``_lmmse`` :215-216."""
    means = np.asarray(means, dtype=complex)
    covs = np.asarray(covs, dtype=complex)
    if A is None:
        A = np.eye(means.shape[1], dtype=complex)
    prep = mofa_prepare(means, covs, A, snr_dB, n_bits, quantizer_type, quantizer)
    logrs = mofa_log_resp_unnorm(y, prep, amps)
    proba = np.exp(logrs - _log_sum(logrs)[None, :]).T
    hk = _component_lmmse(y, means, covs, prep)
    h_est = _combine(proba, hk, n_summands_or_proba, lambda: np.exp(logrs).argmax(axis=0))
    if return_aux:
        return h_est, dict(proba=proba, logrs=logrs, prep=prep)
    return h_est


# --------------------------------------------------------------------------------------
# K = 1 baselines: global / genie Bussgang-LMMSE and Bussgang-LS
# --------------------------------------------------------------------------------------

def toeplitz_cov(t):
    """``toeplitz(t).T`` of the scripts (modules/utils.py:115-150 then ``.T``): first ROW ``t``, first column ``conj(t)``."""
    return scipy.linalg.toeplitz(np.conj(t), t)


def _baseline_operators(C, A, snr_dB, n_bits, quantizer_type, quantizer):
    """``(A_eff, C_r, C_y)`` of one covariance -- estimators/blmmse.py:27-37 (1 bit), :46-57 (b bit), :39-45 (inf)."""
    cy = A @ C @ A.conj().T + 10 ** (-snr_dB / 10) * np.eye(A.shape[0])
    if n_bits == 1:
        psi = np.real(np.diag(1 / np.sqrt(np.diag(cy))))
        return np.sqrt(2 / np.pi) * psi @ A, _quantised_cov(cy, None, 1), cy
    if n_bits == np.inf or n_bits == 'inf':
        return A, cy, cy
    b = _bussgang_diag(np.real(np.diag(cy)), snr_dB, n_bits, quantizer_type, quantizer)
    return np.diag(b) @ A, b[0] ** 2 * cy + (1 - b[0] ** 2) * np.diag(np.diag(cy)), cy      # A_buss[0, 0] (blmmse.py:56)


def blmmse_estimate_global(y, C, snr_dB, A=None, n_bits=1, quantizer_type='uniform', quantizer=None, Cr=None):
    """``BLMMSE.estimate_global`` -- estimators/blmmse.py:61-97: one filter ``C A_eff^H pinv(C_r)`` for all pilots."""
    A = np.eye(y.shape[1], dtype=complex) if A is None else A
    a_eff, cr, _ = _baseline_operators(C, A, snr_dB, n_bits, quantizer_type, quantizer)
    if Cr is not None and n_bits != 1 and n_bits != np.inf:
        cr = Cr
    return y @ (C @ a_eff.conj().T @ np.linalg.pinv(cr)).T


def blmmse_estimate_genie(y, t, snr_dB, A=None, n_bits=1, quantizer_type='uniform', quantizer=None):
    """``BLMMSE.estimate_genie`` -- estimators/blmmse.py:21-58: per-pilot Toeplitz covariance ``toeplitz(t_b).T``."""
    A = np.eye(y.shape[1], dtype=y.dtype) if A is None else A
    out = np.zeros((y.shape[0], A.shape[1]), dtype=complex)
    for b in range(y.shape[0]):
        C = toeplitz_cov(t[b])
        a_eff, cr, _ = _baseline_operators(C, A, snr_dB, n_bits, quantizer_type, quantizer)
        out[b] = C @ a_eff.conj().T @ np.linalg.solve(cr, y[b])
    return out


def ls_estimate_global(y, C, snr_dB, A=None, n_bits=1, quantizer_type='uniform', quantizer=None):
    """``LS.estimate_global`` -- estimators/LS.py:54-74: least squares w.r.t. the Bussgang-effective pilot matrix."""
    A = np.eye(y.shape[1], dtype=complex) if A is None else A
    a_eff, _, _ = _baseline_operators(C, A, snr_dB, n_bits, quantizer_type, quantizer)
    return np.linalg.lstsq(a_eff, y.T, rcond=None)[0].T


def ls_estimate_genie(y, t, snr_dB, A=None, n_bits=1, quantizer_type='uniform', quantizer=None):
    """``LS.estimate_genie`` -- estimators/LS.py:21-52 (finite ``n_bits``; the reference's infinite-resolution branch is broken)."""
    A = np.eye(y.shape[1], dtype=y.dtype) if A is None else A
    out = np.zeros((y.shape[0], A.shape[1]), dtype=complex)
    for b in range(y.shape[0]):
        a_eff, _, _ = _baseline_operators(toeplitz_cov(t[b]), A, snr_dB, n_bits, quantizer_type, quantizer)
        out[b] = np.linalg.lstsq(a_eff, y[b], rcond=None)[0]
    return out


def rate_lower_bound(h_est, h, buss, Cq):
    """Rate lower bound of the scripts -- Bussgang_GMM.py:291-309 (identical in Bussgang_MFA.py:154-172), loop form."""
    res = np.array(h_est, dtype=complex)
    norm_fac = np.clip(np.sum(np.abs(res) ** 2, axis=1), 1e-1, np.inf)
    for i in range(res.shape[0]):
        res[i] /= norm_fac[i]
    inner = np.squeeze(np.expand_dims(res.conj(), 1) @ buss @ np.expand_dims(h, 2))
    num = np.abs(np.mean(inner, axis=0)) ** 2
    den1 = np.var(inner, axis=0)
    den2 = np.real(np.squeeze(np.expand_dims(res.conj(), 1) @ Cq @ np.expand_dims(res, 2)))
    den2 = np.mean(den2, axis=0)
    return float(np.log2(1 + num / (den1 + den2)))


# --------------------------------------------------------------------------------------
# Metric
# --------------------------------------------------------------------------------------

def mse(h_est, h):
    """modules/utils.py:617-618 -- the scripts' "NMSE" (Bussgang_GMM.py:289)."""
    return np.sum(np.abs(h_est - h) ** 2) / h.size


# --------------------------------------------------------------------------------------
# Seeded synthetic parameter / data generators -- shared by tests and bench
# --------------------------------------------------------------------------------------

def random_psd_gmm(K, N, seed=0, mean_scale=0.0):
    """P-rand: ``C = X X^H / (2N)`` scaled to ``tr C = N``; ``w = U(0,1)^K / sum``."""
    rng = np.random.default_rng(seed)
    covs = np.empty((K, N, N), dtype=complex)
    for k in range(K):
        X = crandn(N, 2 * N, rng=rng)
        C = X @ X.conj().T / (2 * N)
        C = C * (N / np.real(np.trace(C)))
        covs[k] = 0.5 * (C + C.conj().T)
    w = rng.random(K)
    w = w / w.sum()
    means = mean_scale * crandn(K, N, rng=rng) if mean_scale else np.zeros((K, N), dtype=complex)
    return means, covs, w


def random_mfa(K, N, M, seed=0, mean_scale=0.0):
    """P-mfa: ``Lambda ~ CN(0, 1/M)``, ``psi ~ 0.02 + 0.1 U(0,1)``."""
    rng = np.random.default_rng(seed)
    lambdas = crandn(K, N, M, rng=rng) / np.sqrt(M)
    psis = 0.02 + 0.1 * rng.random((K, N))
    amps = rng.random(K)
    amps = amps / amps.sum()
    means = mean_scale * crandn(K, N, rng=rng) if mean_scale else np.zeros((K, N), dtype=complex)
    return means, lambdas, psis, amps


def circulant_gmm(K, n1, n2, seed=0):
    """P-circ/BCCB: ``c_k = U(0,1)^N ** 3 + 1e-3`` mean-normalised, ``C_k = F^H diag(c_k) F``,
    ``F = F_n1 (x) F_n2`` unitary (modules/gmm_cplx_bussgang.py:123-125)."""
    rng = np.random.default_rng(seed)
    N = n1 * n2
    c = rng.random((K, N)) ** 3 + 1e-3
    c = c / c.mean(axis=1, keepdims=True)
    F1 = np.fft.fft(np.eye(n1)) / np.sqrt(n1)
    F2 = np.fft.fft(np.eye(n2)) / np.sqrt(n2)
    F = np.kron(F1, F2)
    covs = np.einsum('ji,kj,jl->kil', F.conj(), c, F)
    w = rng.random(K)
    w = w / w.sum()
    return c, covs, w, F


def sample_gmm_channels(means, covs, weights, B, seed=1):
    """``k_b ~ Cat(w)``, ``h_b = mu_k + C_k^{1/2} crandn``; returns ``(h [B,N], noise [B,N], labels)``."""
    rng = np.random.default_rng(seed)
    K, N = means.shape
    lab = rng.choice(K, size=B, p=weights)
    h = np.empty((B, N), dtype=complex)
    for k in np.unique(lab):
        idx = np.nonzero(lab == k)[0]
        L = np.linalg.cholesky(covs[k] + 1e-12 * np.eye(N))
        h[idx] = means[k] + crandn(idx.size, N, rng=rng) @ L.T
    noise = crandn(B, N, rng=rng)
    return h, noise, lab
