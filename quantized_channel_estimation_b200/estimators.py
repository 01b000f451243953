"""The K = 1 baselines of the reference's experiment scripts -- ``BLMMSE`` (estimators/blmmse.py:15-97) and ``LS``
(estimators/LS.py:15-74) with their ``estimate_global`` / ``estimate_genie`` methods and ``mp_eval`` -- on the GPU.

* ``estimate_global`` (one covariance for all pilots) is the degenerate single-component case of the Bussgang-GMM
  estimator: the filter ``C A_eff^H pinv(C_r)`` (or the LS pseudo-inverse) is formed once on the host like the reference
  does, and applied to the batch by the dense estimate kernels of ``libqce_b200.so`` (``DenseModel`` with K = 1, identity
  whitening, zero offsets: ``h = W r``) -- tensor-core path for pilots on a quantiser grid.
* ``estimate_genie`` (one Toeplitz covariance PER PILOT, ``C_b = toeplitz(t_b).T``) has a different shape -- a batch of
  small ``N x N`` solves -- and runs as batched torch / cuSOLVER linear algebra in complex128 (plumbing, not a hand-written
  kernel); the arithmetic follows the reference line by line.

Inputs may be numpy arrays (reference behaviour; the result is a numpy array) or torch CUDA tensors (result stays on the GPU).
"""
import math

import numpy as np
import torch

from . import _lib, engine, precompute


def _is_inf(n_bits):
    return n_bits == 'inf' or n_bits == np.inf


def _device():
    _lib.require_device()
    return torch.device('cuda', torch.cuda.current_device())


def _to_dev(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(dev, torch.complex128)
    return torch.as_tensor(np.asarray(x), dtype=torch.complex128, device=dev)


def _bussgang_operators(C, A, snr_dB, n_bits, quantizer_type, quantizer):
    """``(A_eff, C_r, C_y)`` for a batch of covariances ``C [G, N, N]`` (torch complex128) -- blmmse.py:27-37 (1 bit),
    :46-57 (b bit), :39-45 (infinite resolution)."""
    eye = torch.eye(A.shape[0], dtype=C.dtype, device=C.device)
    Cy = A @ C @ A.conj().T + 10 ** (-snr_dB / 10) * eye
    d = torch.diagonal(Cy, dim1=-2, dim2=-1).real
    if n_bits == 1:
        psi = 1 / torch.sqrt(d)
        A_eff = math.sqrt(2 / math.pi) * psi[..., :, None] * A
        scale = psi[..., :, None] * psi[..., None, :]
        Cr = 2 / math.pi * torch.complex(torch.asin((Cy.real * scale).clamp(-1.0, 1.0)), torch.asin((Cy.imag * scale).clamp(-1.0, 1.0)))
        return A_eff, Cr, Cy
    if _is_inf(n_bits):
        return A.expand(C.shape[:-2] + A.shape), Cy, Cy
    b = torch.as_tensor(precompute.bussgang_gain(d.cpu().numpy(), snr_dB, n_bits, quantizer_type, quantizer), dtype=torch.float64,
                        device=C.device)
    b0 = (b[..., 0] ** 2)[..., None, None]                        # the reference uses A_buss[0, 0] (blmmse.py:56)
    Cr = b0 * Cy + (1 - b0) * torch.diag_embed(torch.diagonal(Cy, dim1=-2, dim2=-1))
    return b[..., :, None] * A, Cr, Cy


def _toeplitz_covs(t):
    """``toeplitz(t_b).T`` for every row of ``t [G, N]``: first row ``t_b``, first column ``conj(t_b)``."""
    N = t.shape[1]
    i = torch.arange(N, device=t.device)
    diff = i[None, :] - i[:, None]                               # j - i
    upper = t[:, diff.clamp(min=0)]
    lower = t.conj()[:, (-diff).clamp(min=0)]
    return torch.where(diff[None] >= 0, upper, lower)


class _Baseline:
    def __init__(self, snr):
        self.snr = snr
        self.rho = 10 ** (0.1 * snr)
        self.sigma2 = 1 / self.rho
        self.precision = 'auto'

    def _apply_filter(self, y, W, n_bits, quantizer_type):
        """``h_b = W r_b`` for the whole batch on the dense estimate kernels (K = 1)."""
        dev = _device()
        W = _to_dev(W, dev)
        n_ant, n_obs = W.shape
        prep = dict(Linv=torch.eye(n_obs, dtype=torch.complex128, device=dev)[None].contiguous(), W=W[None].contiguous(),
                    zoff=torch.zeros((1, n_obs), dtype=torch.complex128, device=dev),
                    hoff=torch.zeros((1, n_ant), dtype=torch.complex128, device=dev),
                    logc=torch.zeros(1, dtype=torch.float64, device=dev),
                    data_scale=precompute.data_scale_for(self.snr, np.inf if _is_inf(n_bits) else n_bits, quantizer_type),
                    n_obs=n_obs, n_ant=n_ant, n_comp=1)
        model = engine.DenseModel(prep)
        if isinstance(y, torch.Tensor) and y.is_cuda:
            return model.estimate(y, 'all', self.precision)
        out = model.estimate_host(y.numpy() if isinstance(y, torch.Tensor) else y, 'all', self.precision)
        return torch.from_numpy(out) if isinstance(y, torch.Tensor) else out

    @staticmethod
    def _finish(h, y):
        if isinstance(y, torch.Tensor):
            return h if y.is_cuda else h.cpu()
        return h.cpu().numpy()


class BLMMSE(_Baseline):
    """Bussgang-LMMSE with a global sample covariance or the per-pilot (genie) covariance -- estimators/blmmse.py:15-97."""

    def estimate_global(self, y, C, A=None, n_bits=1, quantizer_type='uniform', quantizer=None, Cr=None):
        dev = _device()
        Cd = _to_dev(C, dev)
        Ad = torch.eye(y.shape[1], dtype=torch.complex128, device=dev) if A is None else _to_dev(A, dev)
        A_eff, Cr_own, _ = _bussgang_operators(Cd, Ad, self.snr, n_bits, quantizer_type, quantizer)
        if Cr is not None and n_bits != 1 and not _is_inf(n_bits):
            Cr_own = _to_dev(Cr, dev)
        W = Cd @ A_eff.conj().T @ torch.linalg.pinv(Cr_own)
        return self._apply_filter(y, W, n_bits, quantizer_type)

    def estimate_genie(self, y, t, A=None, n_bits=1, quantizer_type='uniform', quantizer=None, Cr=None, chunk=2048):
        dev = _device()
        yd, td = _to_dev(y, dev), _to_dev(t, dev)
        Ad = torch.eye(yd.shape[1], dtype=torch.complex128, device=dev) if A is None else _to_dev(A, dev)
        out = torch.empty((yd.shape[0], Ad.shape[1]), dtype=torch.complex128, device=dev)
        for b0 in range(0, yd.shape[0], chunk):
            C = _toeplitz_covs(td[b0:b0 + chunk])
            A_eff, Cr_b, _ = _bussgang_operators(C, Ad, self.snr, n_bits, quantizer_type, quantizer)
            x = torch.linalg.solve(Cr_b, yd[b0:b0 + chunk, :, None])
            out[b0:b0 + chunk] = (C @ (A_eff.conj().transpose(-1, -2) @ x))[:, :, 0]
        return self._finish(out, y)


class LS(_Baseline):
    """Least squares w.r.t. the Bussgang-effective pilot matrix -- estimators/LS.py:15-74."""

    def estimate_global(self, y, C, A=None, n_bits=1, quantizer_type='uniform', quantizer=None):
        dev = _device()
        Ad = torch.eye(y.shape[1], dtype=torch.complex128, device=dev) if A is None else _to_dev(A, dev)
        A_eff, _, _ = _bussgang_operators(_to_dev(C, dev), Ad, self.snr, n_bits, quantizer_type, quantizer)
        return self._apply_filter(y, torch.linalg.pinv(A_eff), n_bits, quantizer_type)

    def estimate_genie(self, y, t, A=None, n_bits=1, quantizer_type='uniform', quantizer=None, chunk=2048):
        if _is_inf(n_bits):        # independent of the covariance (LS.py:34-36, whose own code assigns the lstsq tuple)
            return self.estimate_global(y, np.eye(np.asarray(t).shape[1]) if not isinstance(t, torch.Tensor) else torch.eye(t.shape[1]),
                                        A, n_bits, quantizer_type, quantizer)
        dev = _device()
        yd, td = _to_dev(y, dev), _to_dev(t, dev)
        Ad = torch.eye(yd.shape[1], dtype=torch.complex128, device=dev) if A is None else _to_dev(A, dev)
        out = torch.empty((yd.shape[0], Ad.shape[1]), dtype=torch.complex128, device=dev)
        for b0 in range(0, yd.shape[0], chunk):
            A_eff, _, _ = _bussgang_operators(_toeplitz_covs(td[b0:b0 + chunk]), Ad, self.snr, n_bits, quantizer_type, quantizer)
            out[b0:b0 + chunk] = (torch.linalg.pinv(A_eff) @ yd[b0:b0 + chunk, :, None])[:, :, 0]
        return self._finish(out, y)


def mp_eval(obj, y, toep, h_true, genie, A=None, n_bits=1, quantizer_type=None, quantizer=None, Cr=None):
    """``mp_eval`` of estimators/blmmse.py:7-12 / estimators/LS.py:7-12 (``h_true`` is unused there as well)."""
    if genie:
        return obj.estimate_genie(y, toep, A, n_bits, quantizer_type, quantizer)
    if isinstance(obj, BLMMSE):
        return obj.estimate_global(y, toep, A, n_bits, quantizer_type, quantizer, Cr)
    return obj.estimate_global(y, toep, A, n_bits, quantizer_type, quantizer)
