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


def prepare(means, covs, weights, A, snr_dB, n_bits=1, quantizer_type='uniform', quantizer=None, device=None):
    """Return the parameter blocks as a dict of contiguous torch tensors on ``device``.

    means [K,N], covs [K,N,N] complex; weights [K]; A [n_obs,N].  Raises ``ValueError`` (reference
    message) if some ``C_r,k`` is not positive definite.
    """
    if device is None:
        device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    cd, fd = torch.complex128, torch.float64
    mu = torch.as_tensor(np.asarray(means), dtype=cd, device=device)
    Ch = torch.as_tensor(np.asarray(covs), dtype=cd, device=device)
    w = torch.as_tensor(np.asarray(weights), dtype=fd, device=device)
    Am = torch.as_tensor(np.asarray(A), dtype=cd, device=device)
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
