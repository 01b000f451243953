"""Golden vectors at tensor-core shapes (N = 16 / 32) from the UNMODIFIED reference: they pin the tcgen05 path directly on reference
outputs (tests/test_gpu_parity.py::test_tc_golden_reference), not only through the oracle.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_tc.py          (build container only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                       # noqa: E402  (puts /root/reference on sys.path, imports its modules)

MODES = {'all': 'all', 'top1': 1, 'top3': 3, 'cum90': 0.9}


def gmm_tc_cases():
    out = {}
    cfgs = [  # tag, K, N, pilots, mean_scale, n_bits, qtype, snr, B
        ('n32_b1_zm', 8, 32, 1, 0.0, 1, 'uniform', 10, 96),       # config-1 antenna count, 1 bit
        ('n16_b2u_mean', 6, 16, 1, 0.3, 2, 'uniform', 5, 128),    # uniform 2 bit (odd-integer grid), non-zero means
        ('n16_b3l_zm', 6, 16, 1, 0.0, 3, 'lloyd', 10, 128),       # Lloyd-Max labels: pilots off the grid -> three-pass path
        ('n16_b1_pilots2', 5, 16, 2, 0.2, 1, 'uniform', 0, 128),  # two pilots: n_obs = 32, n_ant = 16
        ('n16_binf_mean', 5, 16, 1, 0.2, np.inf, 'uniform', 15, 96),   # unquantised observations
    ]
    for i, (tag, K, N, npil, ms, nb, qt, snr, B) in enumerate(cfgs):
        rng = np.random.default_rng(700 + i)
        means, covs, w = mg.rand_gmm(rng, K, N, ms)
        h = mg.sample(rng, means, covs, w, B)
        if npil == 1:
            A = np.eye(N, dtype=complex)
        else:
            x = np.exp(2j * np.pi * rng.random(npil))
            A = np.kron(x[:, None], np.eye(N)).astype(complex)
        noise = mg.crandn(rng, B, A.shape[0])
        qz = (None, None, None) if (nb == 1 or nb == np.inf) else mg.ut.get_quantizer_gauss([snr], nb, qt)[snr]
        r = mg.ref_observation(h, snr, A, nb, qz[0], qz[1], noise)
        out[f'{tag}_means'], out[f'{tag}_covs'], out[f'{tag}_w'] = means, covs, w
        out[f'{tag}_A'], out[f'{tag}_r'] = A, r
        out[f'{tag}_snr'], out[f'{tag}_nbits'], out[f'{tag}_qtype'] = np.asarray(float(snr)), np.asarray(float(nb)), np.asarray(qt)
        if qz[0] is not None:
            out[f'{tag}_thr'], out[f'{tag}_lab'] = qz[0], qz[1]
        for mtag, mode in MODES.items():
            g = mg.Gmm_nbit(n_components=K, covariance_type='full')
            g.params['zero_mean'] = (ms == 0.0)
            g.means_cplx, g.covs_cplx = means.copy(), covs.copy()
            g.gm.weights_ = w.copy()
            out[f'{tag}_est_{mtag}'] = g.estimate_from_y(r, snr, N, A=A, n_summands_or_proba=mode, n_bits=nb, quantizer_type=qt, quantizer=qz)
            if mtag == 'all':
                out[f'{tag}_wlp'] = g._estimate_weighted_log_prob(r)
    return out


if __name__ == '__main__':
    d = gmm_tc_cases()
    path = os.path.join(HERE, 'gmm_tc.npz')
    np.savez_compressed(path, **d)
    print('gmm_tc', len(d), 'arrays', os.path.getsize(path), 'bytes')
