import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f'{name}.npz'), allow_pickle=False))


@pytest.fixture(scope='session')
def golden_quantizer():
    return load_golden('quantizer')


@pytest.fixture(scope='session')
def golden_quant():
    return load_golden('quant')


@pytest.fixture(scope='session')
def golden_gmm():
    return load_golden('gmm')


@pytest.fixture(scope='session')
def golden_gmm_tc():
    return load_golden('gmm_tc')


@pytest.fixture(scope='session')
def golden_mfa():
    return load_golden('mfa')


@pytest.fixture(scope='session')
def golden_baselines():
    return load_golden('baselines')


BASELINE_TAGS = [f'{b}_{a}' for b in ('b1', 'u2', 'l3', 'inf') for a in ('I', 'A2')]


def baseline_case(g, tag):
    """(r, A, n_bits, quantizer_type, quantizer) of one fixture case of tests/golden/make_golden_baselines.py."""
    import numpy as np
    nb = float(g[tag + '_nbits'])
    nb = int(nb) if np.isfinite(nb) else np.inf
    N = g['C_glob'].shape[0]
    A = None if tag.endswith('_I') else np.kron(np.array([[1.0], [1j]]), np.eye(N))
    qz = (g[tag + '_thr'], g[tag + '_lab'], None) if tag + '_thr' in g else (None, None, None)
    return g[tag + '_r'], A, nb, str(g[tag + '_qtype']), qz


GMM_TAGS = ['b1_zm', 'b1_mean', 'b2u_mean', 'b3l_zm', 'binf_mean', 'b1_pilots2', 'b2u_pilots2', 'b1_k1']
GMM_MODES = {'all': 'all', 'top1': 1, 'top3': 3, 'cum90': 0.9}
# tests/golden/make_golden_tc.py: reference outputs at shapes the tensor-core kernels are instantiated for
GMM_TC_TAGS = ['n32_b1_zm', 'n16_b2u_mean', 'n16_b3l_zm', 'n16_b1_pilots2', 'n16_binf_mean']
MFA_TAGS = ['b1_zm', 'b2u_mean', 'b3l_mean']
MFA_MODES = {'all': 'all', 'top1': 1, 'top2': 2, 'cum90': 0.9}


def golden_quantizer_tuple(g, tag):
    if f'{tag}_thr' in g:
        return (g[f'{tag}_thr'], g[f'{tag}_lab'], None)
    return (None, None, None)


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
