"""Model files as the reference scripts save them (``joblib.dump(model, '...sav')``, Bussgang_GMM.py:262-264 /
Bussgang_MFA.py:125-127), made from UNMODIFIED reference classes with injected seeded parameters.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sav.py
"""
import os
import sys

import joblib
import numpy as np

REF = os.environ.get("QCE_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from modules.gmm_cplx_bussgang import Gmm_nbit                # noqa: E402
from modules.mofa_cplx_bussgang import Mofa                   # noqa: E402


def crandn(rng, *shape):
    return np.sqrt(0.5) * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def main():
    rng = np.random.default_rng(7)
    K, N, M = 3, 8, 2
    covs = np.empty((K, N, N), dtype=complex)
    for k in range(K):
        X = crandn(rng, N, 2 * N)
        covs[k] = X @ X.conj().T / (2 * N)
    w = rng.random(K)
    w /= w.sum()
    gmm = Gmm_nbit(n_components=K, covariance_type='full')
    gmm.params['zero_mean'] = False
    gmm.means_cplx = 0.1 * crandn(rng, K, N)
    gmm.covs_cplx = covs
    gmm.gm.weights_ = w
    joblib.dump(gmm, os.path.join(HERE, 'ref_gmm.sav'))
    mfa = Mofa(K, M, verbose=False)
    mfa.D = N
    mfa.means = 0.1 * crandn(rng, K, N)
    mfa.lambdas = crandn(rng, K, N, M) / np.sqrt(M)
    mfa.psis = 0.02 + 0.1 * rng.random((K, N))
    mfa.amps = w
    mfa.covs = mfa.lambdas @ np.transpose(mfa.lambdas.conj(), [0, 2, 1]) + np.stack([np.diag(p) for p in mfa.psis])
    joblib.dump(mfa, os.path.join(HERE, 'ref_mofa.sav'))
    np.savez(os.path.join(HERE, 'ref_sav_params.npz'), gmm_means=gmm.means_cplx, gmm_covs=covs, w=w, mfa_means=mfa.means,
             mfa_lambdas=mfa.lambdas, mfa_psis=mfa.psis, mfa_covs=mfa.covs)
    print('wrote ref_gmm.sav, ref_mofa.sav, ref_sav_params.npz')


if __name__ == '__main__':
    main()
