"""Golden vectors for the K = 1 baselines (global / genie Bussgang-LMMSE and Bussgang-LS) from the UNMODIFIED reference
(estimators/blmmse.py, estimators/LS.py).  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_baselines.py
"""
import os
import sys

import numpy as np

REF = os.environ.get("QCE_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import modules.utils as ut                                    # noqa: E402
from estimators.blmmse import BLMMSE                          # noqa: E402
from estimators.LS import LS                                  # noqa: E402


def crandn(rng, *shape):
    return np.sqrt(0.5) * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def main():
    rng = np.random.default_rng(2024)
    N, B, snr = 16, 24, 5
    # per-sample Toeplitz covariances from a few spectral lines (positive definite), unit diagonal
    t = np.zeros((B, N), dtype=complex)
    for b in range(B):
        f, p = rng.random(4), rng.random(4) + 0.1
        p = 0.9 * p / p.sum()
        t[b] = (p[None, :] * np.exp(2j * np.pi * f[None, :] * np.arange(N)[:, None])).sum(1)
        t[b, 0] += 0.1
    covs = np.stack([ut.toeplitz(t[b]).T for b in range(B)])
    h = np.stack([np.linalg.cholesky(covs[b]) @ crandn(rng, N) for b in range(B)])
    C_glob = covs.mean(axis=0)
    out = dict(t=t, h=h, C_glob=C_glob, snr=float(snr))
    A2 = np.kron(np.array([[1.0], [1j]]), np.eye(N))
    cases = [('b1', 1, 'uniform'), ('u2', 2, 'uniform'), ('l3', 3, 'lloyd'), ('inf', np.inf, 'uniform')]
    for tag, nb, qt in cases:
        qz = ut.get_quantizer_gauss([snr], nb, qt)[snr] if np.isfinite(nb) and nb > 1 else (None, None, None)
        for atag, A in (('I', None), ('A2', A2)):
            Am = np.eye(N) if A is None else A
            noise = crandn(rng, B, Am.shape[0])
            y = h @ Am.T + 10 ** (-snr / 20) * noise
            r = y if not np.isfinite(nb) else ut.quant(y, nb, qz[0], qz[1])
            key = f'{tag}_{atag}'
            out[key + '_r'] = r
            out[key + '_nbits'] = float(nb)
            out[key + '_qtype'] = qt
            if qz[0] is not None:
                out[key + '_thr'], out[key + '_lab'] = qz[0], qz[1]
            out[key + '_blmmse_global'] = BLMMSE(snr).estimate_global(r, C_glob, A, nb, qt, qz)
            out[key + '_ls_global'] = LS(snr).estimate_global(r, C_glob, A, nb, qt, qz)
            out[key + '_blmmse_genie'] = BLMMSE(snr).estimate_genie(r, t, A, nb, qt, qz)
            if np.isfinite(nb):      # the reference's LS.estimate_genie is broken for n_bits = inf (assigns the lstsq tuple)
                out[key + '_ls_genie'] = LS(snr).estimate_genie(r, t, A, nb, qt, qz)
    np.savez_compressed(os.path.join(HERE, 'baselines.npz'), **out)
    print('wrote baselines.npz with', len(out), 'arrays')


if __name__ == '__main__':
    main()
