"""Golden numbers for fit(): run the UNMODIFIED reference EM (with the 4-line test-side compat shim of SURVEY.md section 8c
for the installed numpy / sklearn) on seeded synthetic data and store the data-generating parameters, the seeds and the
average log-likelihood the reference's fitted model reaches.  EM depends on k-means / random initialisation, so the
comparison in tests/test_fit_cpu.py is statistical (our fit must reach the same likelihood level), not bit-wise.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fit.py
"""
import os
import sys

import numpy as np

REF = os.environ.get('QCE_REFERENCE', '/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

np.infty = np.Inf = np.inf                                      # numpy >= 2 dropped the aliases the reference uses
from sklearn.mixture import GaussianMixture                     # noqa: E402
GaussianMixture._check_n_features = lambda s, X, reset=True: setattr(s, 'n_features_in_', X.shape[1])
GaussianMixture._print_verbose_msg_init_end = lambda s, *a, **k: None

from modules.gmm_cplx_bussgang import Gmm_nbit                  # noqa: E402
from modules.mofa_cplx_bussgang import Mofa                     # noqa: E402
from fit_common import avg_loglik, make_data                   # noqa: E402

out = {}
for tag, ctype, blocks, zm in [('full_zm', 'full', None, True), ('full_mean', 'full', None, False),
                               ('circ', 'circulant', None, True), ('bccb', 'block-circulant', (2, 4), True),
                               ('toep', 'toeplitz', None, True), ('btoep', 'block-toeplitz', (2, 4), True)]:
    h, true = make_data(tag)
    g = Gmm_nbit(n_components=3, covariance_type=ctype, random_state=0, max_iter=200, tol=1e-5)
    g.fit(h, blocks=blocks, zero_mean=zm)
    out[f'{tag}_ref_ll'] = np.asarray(avg_loglik(h, g.gm.weights_, g.means_cplx, g.covs_cplx))
    out[f'{tag}_true_ll'] = np.asarray(avg_loglik(h, *true))
    print(tag, out[f'{tag}_ref_ll'], out[f'{tag}_true_ll'])
h, true = make_data('mfa')
np.random.seed(0)
m = Mofa(3, 2, verbose=False, maxiter=200)
m.fit(h, zero_mean=False)
out['mfa_ref_ll'] = np.asarray(avg_loglik(h, m.amps, m.means, m.covs))
out['mfa_true_ll'] = np.asarray(avg_loglik(h, *true))
print('mfa', out['mfa_ref_ll'], out['mfa_true_ll'])
np.savez_compressed(os.path.join(HERE, 'fit.npz'), **out)
