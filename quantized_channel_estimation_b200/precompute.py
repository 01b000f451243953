"""Per-component host precompute for one (SNR, bit width, quantiser) setting.

Restates ``Gmm_nbit._prepare_for_prediction`` (reference modules/gmm_cplx_bussgang.py:246-328) and
``Mofa._prepare_for_prediction`` (modules/mofa_cplx_bussgang.py:162-212) batched over the K
components in torch float64 / complex128 (on the GPU when there is one -- this is O(K N^3) setup
work done once per SNR, not the per-sample hot path), and folds the results into the parameter
blocks the kernels consume (include/qce_b200.h, qce_model_set_params):

    Linv_k = L_k^-1 with C_r,k = L_k L_k^H        zoff_k = Linv_k m_r,k
    W_k    = C_h,k A_eff,k^H C_r,k^-1              hoff_k = mu_k - W_k m_r,k
    logc_k = ln w_k - n_obs ln(pi) - ln|C_r,k|
"""
import math

import numpy as np
import torch

from . import lloyd_max_quantizer as quant_lloyd
from . import uniform_quantizer as quant_uni

NOT_PD_MSG = ("Fitting the mixture model failed because some components have "
              "ill-defined empirical covariance (for instance caused by singleton "
              "or collapsed samples). Try to decrease the number of components, "
              "or increase reg_covar.")          # the reference's ValueError text (gmm:33-37)


def _is_inf(n_bits):
    return n_bits == 'inf' or n_bits == np.inf


def bussgang_gain(var, snr_dB, n_bits, quantizer_type, quantizer):
    """Per-antenna Bussgang gain ``b [K, n_obs]`` from per-antenna variances (numpy, float64).
    Selection logic of gmm:274-284 / mofa:171-180."""
    if n_bits == 1:
        return math.sqrt(2 / math.pi) * (1 / np.sqrt(var))
    if _is_inf(n_bits):
        return np.ones_like(var)
    if quantizer_type == 'uniform':
        return quant_uni.bussgang_diag(snr_dB, n_bits, var)
    if quantizer_type == 'lloyd':
        return quant_lloyd.bussgang_diag(n_bits, var, quantizer)
    raise NotImplementedError(f'Quantizer type {quantizer_type} not implemented!')


def data_scale_for(snr_dB, n_bits, quantizer_type):
    """Grid on which every real/imaginary part of r lies (0 = arbitrary reals)."""
    if n_bits == 1:
        return float(1 / np.sqrt(2))
    if _is_inf(n_bits) or quantizer_type != 'uniform' or n_bits > 8:
        return 0.0
    return float(quant_uni.get_uniform_quant_step(snr_dB, n_bits)) / 2      # labels are odd multiples of step/2


def _t(x, dtype, device):
    """numpy -> torch on ``device`` (the model's parameter arrays are read-only: torch warns when it wraps those without a copy)."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', UserWarning)
        return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


def prepare(means, covs, weights, A, snr_dB, n_bits=1, quantizer_type='uniform', quantizer=None, device=None):
    """Return the parameter blocks as a dict of contiguous torch tensors on ``device``.

    means [K,N], covs [K,N,N] complex; weights [K]; A [n_obs,N].  Raises ``ValueError`` (reference
    message) if some ``C_r,k`` is not positive definite.
    """
    if device is None:
        device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    cd, fd = torch.complex128, torch.float64
    mu = _t(means, cd, device)
    Ch = _t(covs, cd, device)
    w = _t(weights, fd, device)
    Am = _t(A, cd, device)
    K, N = mu.shape
    No = Am.shape[0]
    sigma2 = 10 ** (-snr_dB / 10)

    m_y = mu @ Am.T                                                   # A mu_k              (gmm:256)
    ChAh = Ch @ Am.conj().T                                           # C_h A^H   [K,N,No]
    Cy = Am @ ChAh                                                    # A C_h A^H           (gmm:268)
    eye = torch.eye(No, dtype=cd, device=device)
    Cy = Cy + sigma2 * eye                                            #                     (gmm:269-271)
    var = torch.diagonal(Cy, dim1=1, dim2=2).real                     # [K,No]

    b = torch.as_tensor(bussgang_gain(var.cpu().numpy(), snr_dB, n_bits, quantizer_type, quantizer),
                        dtype=fd, device=device)                      #                     (gmm:274-284)
    m_r = b * m_y                                                     #                     (gmm:287-288)

    if n_bits == 1:                                                   # arcsine law         (gmm:292-301)
        s = 1 / torch.sqrt(var)
        rho = s[:, :, None] * Cy * s[:, None, :]
        Cr = (2 / math.pi) * torch.complex(torch.asin(rho.real.clamp(-1.0, 1.0)), torch.asin(rho.imag.clamp(-1.0, 1.0)))
    elif _is_inf(n_bits):                                             #                     (gmm:302-303)
        Cr = Cy
    else:                                                             # scalar-beta model   (gmm:304-307)
        beta = b.mean(dim=1).clamp(0, 1)
        b2 = (beta ** 2)[:, None, None]
        Cr = b2 * Cy + (1 - b2) * torch.diag_embed(torch.diagonal(Cy, dim1=1, dim2=2))

    L, info = torch.linalg.cholesky_ex(Cr)                            # C_r = L L^H         (gmm:15-47)
    if int(info.max()) != 0 or not bool(torch.isfinite(L.real).all()):
        raise ValueError(NOT_PD_MSG)
    Linv = torch.linalg.solve_triangular(L, eye.expand(K, No, No), upper=False)
    logdet = 2 * torch.log(torch.diagonal(L, dim1=1, dim2=2).real).sum(dim=1)
    Cr_inv = Linv.conj().transpose(1, 2) @ Linv                       # == pinv(C_r) for PD C_r (gmm:321-323)
    W = (ChAh * b[:, None, :]) @ Cr_inv                               # C_h A_eff^H C_r^-1  (gmm:226-228, :326)
    zoff = (Linv @ m_r[:, :, None])[:, :, 0]
    hoff = mu - (W @ m_r[:, :, None])[:, :, 0]
    logc = torch.log(w) - No * math.log(math.pi) - logdet             #                     (gmm:380-386, :435)
    return dict(Linv=Linv.contiguous(), W=W.contiguous(), zoff=zoff.contiguous(), hoff=hoff.contiguous(),
                logc=logc.contiguous(), data_scale=data_scale_for(snr_dB, n_bits, quantizer_type),
                m_r=m_r, C_r=Cr, b=b, n_obs=No, n_ant=N, n_comp=K)


# --------------------------------------------------------------------------------------------------------------
# circulant / block-circulant covariances: everything is diagonal in the (2-D) DFT basis
# --------------------------------------------------------------------------------------------------------------

def dft_matrix(n1, n2):
    """Unitary ``F = F_n1 (x) F_n2`` (reference gmm:105, :123-125)."""
    F1 = np.fft.fft(np.eye(n1)) / np.sqrt(n1)
    F2 = np.fft.fft(np.eye(n2)) / np.sqrt(n2)
    return np.kron(F1, F2)


def detect_blocks(covs, tol=1e-9):
    """Return ``(n1, n2), c [K,N]`` if every ``covs[k] = F^H diag(c_k) F`` for some factorisation ``N = n1 n2``
    (plain circulant = ``(1, N)`` is tried first), else ``(None, None)``."""
    covs = np.asarray(covs)
    N = covs.shape[-1]
    scale = np.linalg.norm(covs)
    cands = [(1, N)] + [(a, N // a) for a in range(2, N) if N % a == 0]
    for n1, n2 in cands:
        F = dft_matrix(n1, n2)
        D = F @ covs @ F.conj().T
        off = D - np.stack([np.diag(np.diag(d)) for d in D])
        if np.linalg.norm(off) <= tol * scale:
            return (n1, n2), np.real(np.diagonal(D, axis1=1, axis2=2)).copy()
    return None, None


def prepare_circulant(c, weights, blocks, snr_dB, n_bits=1, quantizer_type='uniform', quantizer=None, device=None):
    """DFT-domain parameter blocks for ``C_h,k = F^H diag(c_k) F``, ``A = I``, zero means (SURVEY.md section 2.1,
    "structured covariances"): ``inv_lambda_t [N,K]``, ``gain [K,N]``, ``logc [K]`` as float64 torch tensors."""
    if device is None:
        device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    n1, n2 = blocks
    c = np.asarray(c, dtype=float)
    K, N = c.shape
    sigma2 = 10 ** (-snr_dB / 10)
    cy = c + sigma2                                           # eigenvalues of C_y
    d = cy.mean(axis=1)                                       # its constant diagonal
    b = bussgang_gain(d[:, None], snr_dB, n_bits, quantizer_type, quantizer)[:, 0]     # scalar gain per component
    if n_bits == 1:
        col0 = np.fft.ifft2(cy.reshape(K, n1, n2), axes=(1, 2))          # first column of C_y (times 1: F e_0 = 1/sqrt(N))
        rho = col0 / d[:, None, None]
        cr0 = 2 / np.pi * (np.arcsin(np.clip(rho.real, -1.0, 1.0)) + 1j * np.arcsin(np.clip(rho.imag, -1.0, 1.0)))
        lam = np.real(np.fft.fft2(cr0, axes=(1, 2))).reshape(K, N)       # eigenvalues of the (block-)circulant C_r
    elif _is_inf(n_bits):
        lam = cy
    else:
        beta = np.clip(b, 0, 1)
        lam = beta[:, None] ** 2 * cy + (1 - beta[:, None] ** 2) * d[:, None]
    if not np.all(lam > 0):
        raise ValueError(NOT_PD_MSG)
    gain = b[:, None] * c / lam
    logc = np.log(np.asarray(weights, dtype=float)) - N * math.log(math.pi) - np.log(lam).sum(axis=1)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=device)
    return dict(inv_lambda_t=t((1 / lam).T), gain=t(gain), logc=t(logc), n1=n1, n2=n2, n_ant=N, n_comp=K)


# --------------------------------------------------------------------------------------------------------------
# mixture of factor analysers: Woodbury (low-rank + diagonal) form, A = I, n_bits > 1 (or infinite resolution)
# --------------------------------------------------------------------------------------------------------------

def prepare_mfa_woodbury(means, lambdas, psis, amps, snr_dB, n_bits, quantizer_type='uniform', quantizer=None, device=None):
    """Parameter blocks of ``qce_mfa_model_set_params`` (include/qce_b200.h).  ``C_r,k = beta^2 Lambda Lambda^H + Delta_k``
    (reference mofa:199-202 written out for ``C_h = Lambda Lambda^H + diag(psi)``) is never formed."""
    if n_bits == 1:
        raise ValueError('the arcsine law destroys the low-rank structure: 1-bit MFA uses the dense path')
    if device is None:
        device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    cd, fd = torch.complex128, torch.float64
    mu = _t(means, cd, device)
    Lam = _t(lambdas, cd, device)               # [K,N,M]
    psi = _t(psis, fd, device)                  # [K,N]
    w = _t(amps, fd, device)
    K, N, M = Lam.shape
    sigma2 = 10 ** (-snr_dB / 10)
    d = (Lam.real ** 2 + Lam.imag ** 2).sum(-1) + psi + sigma2                        # diag C_y            (mofa:167-169)
    b = torch.as_tensor(bussgang_gain(d.cpu().numpy(), snr_dB, n_bits, quantizer_type, quantizer), dtype=fd, device=device)
    m_r = b * mu                                                                      #                     (mofa:183-184)
    if _is_inf(n_bits):
        beta = torch.ones(K, dtype=fd, device=device)
    else:
        beta = b.mean(dim=1).clamp(0, 1)                                              #                     (mofa:199-202)
    b2 = (beta ** 2)[:, None]
    Delta = b2 * (psi + sigma2) + (1 - b2) * d                                        # diagonal part of C_r
    if not bool((Delta > 0).all()):
        raise ValueError(NOT_PD_MSG)
    invD = 1 / Delta
    U = beta[:, None, None] * Lam                                                     # C_r = U U^H + Delta
    UhD = U.conj().transpose(1, 2) * invD[:, None, :]                                 # U^H Delta^-1       [K,M,N]
    S = torch.eye(M, dtype=cd, device=device) + UhD @ U                               # [K,M,M]
    LS, info = torch.linalg.cholesky_ex(S)
    if int(info.max()) != 0:
        raise ValueError(NOT_PD_MSG)
    T = torch.linalg.solve_triangular(LS, UhD, upper=False)                           # L_S^-1 U^H Delta^-1
    Q = torch.linalg.solve_triangular(LS.conj().transpose(1, 2), T, upper=True)       # S^-1 U^H Delta^-1
    P = Lam.conj().transpose(1, 2) * (b * invD)[:, None, :]                           # Lambda^H B Delta^-1 [K,M,N]
    V1 = P - (P @ U) @ Q
    e = psi * b * invD                                                                # diagonal part of W
    LSinvH = torch.linalg.solve_triangular(LS, torch.eye(M, dtype=cd, device=device).expand(K, M, M), upper=False).conj().transpose(1, 2)
    Y = torch.cat([Lam, -(e[:, :, None] * U) @ LSinvH], dim=2)                        # [K,N,2M]
    D = torch.cat([V1, T], dim=1)                                                     # [K,2M,N]
    logdet = torch.log(Delta).sum(1) + 2 * torch.log(torch.diagonal(LS, dim1=1, dim2=2).real).sum(1)
    logc = torch.log(w) - N * math.log(math.pi) - logdet
    return dict(inv_delta=invD.contiguous(), evec=e.contiguous(), D=D.contiguous(), Y=Y.contiguous(), m_r=m_r.contiguous(),
                mu=mu.contiguous(), logc=logc.contiguous(), n_ant=N, latent=M, n_comp=K)
