"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (the reference cannot travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``/root/reference`` read-only, injects seeded parameters into the
reference estimators (``Gmm_nbit`` / ``Mofa``) exactly the way a fitted model
holds them, calls the reference's own public functions and
stores inputs + outputs as small ``.npz`` fixtures next to this script.  The
fixtures pin ``oracle/qce_oracle.py`` (tests/test_oracle_golden.py) and, on the GPU
box, the CUDA path (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np

REF = os.environ.get("QCE_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import modules.utils as ut                                    # noqa: E402
import modules.uniform_quantizer as quant_uni                 # noqa: E402
import modules.lloyd_max_quantizer as quant_lloyd             # noqa: E402
from modules.gmm_cplx_bussgang import Gmm_nbit                # noqa: E402
from modules.mofa_cplx_bussgang import Mofa                   # noqa: E402


def crandn(rng, *shape):
    return np.sqrt(0.5) * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def rand_gmm(rng, K, N, mean_scale):
    covs = np.empty((K, N, N), dtype=complex)
    for k in range(K):
        X = crandn(rng, N, 2 * N)
        C = X @ X.conj().T / (2 * N)
        C *= N / np.real(np.trace(C))
        covs[k] = 0.5 * (C + C.conj().T)
    w = rng.random(K)
    w /= w.sum()
    means = mean_scale * crandn(rng, K, N)
    return means, covs, w


def sample(rng, means, covs, w, B):
    K, N = means.shape
    lab = rng.choice(K, size=B, p=w)
    h = np.empty((B, N), dtype=complex)
    for b in range(B):
        L = np.linalg.cholesky(covs[lab[b]])
        h[b] = means[lab[b]] + L @ crandn(rng, N)
    return h


def ref_observation(h, snr, A, n_bits, thr, labels, noise):
    """``ut.get_observation_nbit`` with the module-global noise draw replaced by ``noise``."""
    saved = ut.crandn
    ut.crandn = lambda *shape: noise.copy()
    try:
        return ut.get_observation_nbit(h[:, None, :] if False else h, snr, A=A, n_bits=n_bits,
                                       thresholds=thr, cluster=labels)
    finally:
        ut.crandn = saved


def quantizer_cases():
    out = {}
    snrs = [-10, 0, 10, 20]
    for nb in (2, 3, 4):
        q = ut.get_quantizer_gauss(snrs, nb, 'uniform')
        for s in snrs:
            out[f'uni_b{nb}_s{s}_thr'] = q[s][0]
            out[f'uni_b{nb}_s{s}_lab'] = q[s][1]
    for nb in (2, 3):
        q = ut.get_quantizer_gauss([0, 10], nb, 'lloyd')
        for s in (0, 10):
            out[f'lloyd_b{nb}_s{s}_thr'] = q[s][0]
            out[f'lloyd_b{nb}_s{s}_lab'] = q[s][1]
            out[f'lloyd_b{nb}_s{s}_rho'] = np.asarray(q[s][2])
    for nb in range(1, 11):
        out[f'step_b{nb}'] = np.asarray(quant_uni.standard_quantization_step(nb))
        out[f'rhofac_b{nb}'] = np.asarray(quant_uni.standard_distortion_fac(nb))
        out[f'qstep_b{nb}_s5'] = np.asarray(quant_uni.get_uniform_quant_step(5, nb))
        out[f'rhouni_b{nb}_s5'] = np.asarray(quant_uni.get_rho_uniform(5, nb))
        out[f'rholloyd_b{nb}'] = np.asarray(quant_lloyd.get_rho_lloyd(5, nb))
    # Bussgang matrices / quantised variance / C_r on a seeded covariance
    rng = np.random.default_rng(7)
    _, covs, _ = rand_gmm(rng, 1, 6, 0.0)
    Cy = covs[0] + 0.1 * np.eye(6)
    out['buss_Cy'] = Cy
    for nb in (1, 2, 3):
        out[f'buss_uni_b{nb}'] = quant_uni.get_Bussgang_matrix(snr_dB=10, n_bits=nb, Cy=Cy)
        qz = ut.get_quantizer_gauss([10], nb, 'uniform')[10] if nb > 1 else (None, None, None)
        out[f'Cr_uni_b{nb}'] = quant_uni.get_Cr(Cy, n_bits=nb, snr=10, quantizer=qz)
    ql = ut.get_quantizer_gauss([10], 3, 'lloyd')[10]
    out['buss_lloyd_b3'] = quant_lloyd.get_Bussgang_matrix(n_bits=3, Cy=Cy, quantizer=ql)
    qu = ut.get_quantizer_gauss([10], 2, 'uniform')[10]
    out['qvar_in'] = np.array([0.3, 1.0, 2.5])
    out['qvar_uni_b2'] = quant_uni.get_quantized_variance(np.array([0.3, 1.0, 2.5]), qu)
    return out


def quant_cases():
    out = {}
    rng = np.random.default_rng(11)
    y = crandn(rng, 9, 5)
    y[0, 0] = 0.0 + 0.0j
    y[0, 1] = -0.0 + 1.0j
    y[1, 0] = complex(np.nan, -1.0)
    out['y'] = y
    out['q1'] = ut.quant(y, 1)
    qu = ut.get_quantizer_gauss([10], 2, 'uniform')[10]
    yu = y.copy()
    yu[2, 0] = qu[0][0] + 1j * qu[0][2]            # exactly on thresholds
    out['yu'] = yu
    out['q2u'] = ut.quant(yu, 2, qu[0], qu[1])
    ql = ut.get_quantizer_gauss([10], 3, 'lloyd')[10]
    out['q3l'] = ut.quant(y, 3, ql[0], ql[1])
    out['q3l_thr'], out['q3l_lab'] = ql[0], ql[1]
    # observation synthesis with a known noise draw, complex64 channels as SCMMulti returns
    h = crandn(rng, 12, 6).astype(np.complex64)
    noise = crandn(rng, 12, 6)
    out['obs_h'], out['obs_noise'] = h, noise
    for snr in (-5, 10):
        out[f'obs1_s{snr}'] = ref_observation(h, snr, None, 1, None, None, noise)
        qs = ut.get_quantizer_gauss([snr], 2, 'uniform')[snr]
        out[f'obs2u_s{snr}'] = ref_observation(h, snr, None, 2, qs[0], qs[1], noise)
        out[f'obsinf_s{snr}'] = ref_observation(h, snr, None, np.inf, None, None, noise)
    return out


def gmm_cases():
    out = {}
    modes = {'all': 'all', 'top1': 1, 'top3': 3, 'cum90': 0.9}
    cfgs = [  # tag, K, N, pilots, mean_scale, n_bits, qtype, snr
        ('b1_zm', 5, 8, 1, 0.0, 1, 'uniform', 5),
        ('b1_mean', 5, 8, 1, 0.3, 1, 'uniform', 15),
        ('b2u_mean', 4, 8, 1, 0.3, 2, 'uniform', 10),
        ('b3l_zm', 4, 8, 1, 0.0, 3, 'lloyd', 10),
        ('binf_mean', 4, 8, 1, 0.3, np.inf, 'uniform', 10),
        ('b1_pilots2', 4, 6, 2, 0.2, 1, 'uniform', 0),
        ('b2u_pilots2', 4, 6, 2, 0.2, 2, 'uniform', 10),
        ('b1_k1', 1, 8, 1, 0.3, 1, 'uniform', 5),
    ]
    for i, (tag, K, N, npil, ms, nb, qt, snr) in enumerate(cfgs):
        rng = np.random.default_rng(100 + i)
        means, covs, w = rand_gmm(rng, K, N, ms)
        B = 24
        h = sample(rng, means, covs, w, B)
        if npil == 1:
            A = np.eye(N, dtype=complex)
        else:
            x = np.exp(2j * np.pi * rng.random(npil)) / 1.0
            A = np.kron(x[:, None], np.eye(N)).astype(complex)     # utils.get_pilot_matrix semantics (:366)
        noise = crandn(rng, B, A.shape[0])
        if nb == 1 or nb == np.inf:
            qz = (None, None, None)
        else:
            qz = ut.get_quantizer_gauss([snr], nb, qt)[snr]
        r = ref_observation(h, snr, A, nb, qz[0], qz[1], noise)
        out[f'{tag}_means'], out[f'{tag}_covs'], out[f'{tag}_w'] = means, covs, w
        out[f'{tag}_A'], out[f'{tag}_h'], out[f'{tag}_noise'], out[f'{tag}_r'] = A, h, noise, r
        out[f'{tag}_snr'] = np.asarray(float(snr))
        out[f'{tag}_nbits'] = np.asarray(float(nb))
        out[f'{tag}_qtype'] = np.asarray(qt)
        if qz[0] is not None:
            out[f'{tag}_thr'], out[f'{tag}_lab'] = qz[0], qz[1]
        for mtag, mode in modes.items():
            if K == 1 and mtag == 'top3':
                continue
            g = Gmm_nbit(n_components=K, covariance_type='full')
            g.params['zero_mean'] = (ms == 0.0)
            g.means_cplx, g.covs_cplx = means.copy(), covs.copy()
            g.gm.weights_ = w.copy()
            est = g.estimate_from_y(r, snr, N, A=A, n_summands_or_proba=mode, n_bits=nb,
                                    quantizer_type=qt, quantizer=qz)
            out[f'{tag}_est_{mtag}'] = est
            if mtag == 'all':
                out[f'{tag}_proba'] = g.predict_proba_cplx(r)
                out[f'{tag}_wlp'] = g._estimate_weighted_log_prob(r)
                out[f'{tag}_mr'] = g.gm.means_.copy()
                out[f'{tag}_Cr'] = g.gm.covariances_.copy()
    return out


def mfa_cases():
    out = {}
    modes = {'all': 'all', 'top1': 1, 'top2': 2, 'cum90': 0.9}
    cfgs = [('b1_zm', 4, 8, 2, 0.0, 1, 'uniform', 5),
            ('b2u_mean', 4, 8, 2, 0.3, 2, 'uniform', 10),
            ('b3l_mean', 3, 8, 3, 0.3, 3, 'lloyd', 10)]
    for i, (tag, K, N, M, ms, nb, qt, snr) in enumerate(cfgs):
        rng = np.random.default_rng(200 + i)
        lambdas = crandn(rng, K, N, M) / np.sqrt(M)
        psis = 0.02 + 0.1 * rng.random((K, N))
        amps = rng.random(K)
        amps /= amps.sum()
        means = ms * crandn(rng, K, N)
        covs = lambdas @ np.transpose(lambdas.conj(), [0, 2, 1]) + np.stack([np.diag(p) for p in psis])
        B = 20
        h = sample(rng, means, covs, amps, B)
        A = np.eye(N, dtype=complex)
        noise = crandn(rng, B, N)
        qz = (None, None, None) if nb == 1 else ut.get_quantizer_gauss([snr], nb, qt)[snr]
        r = ref_observation(h, snr, A, nb, qz[0], qz[1], noise)
        for nm, v in dict(means=means, lambdas=lambdas, psis=psis, amps=amps, covs=covs, h=h, noise=noise,
                          r=r, snr=np.asarray(float(snr)), nbits=np.asarray(float(nb)),
                          qtype=np.asarray(qt)).items():
            out[f'{tag}_{nm}'] = v
        if qz[0] is not None:
            out[f'{tag}_thr'], out[f'{tag}_lab'] = qz[0], qz[1]
        for mtag, mode in modes.items():
            m = Mofa(n_components=K, latent_dim=M, verbose=False)
            m.D = N
            m.means, m.covs, m.lambdas, m.psis, m.amps = means.copy(), covs.copy(), lambdas.copy(), psis.copy(), amps.copy()
            est = m.estimate_from_y(r, snr, A=A, n_summands_or_proba=mode, n_bits=nb, quantizer_type=qt, quantizer=qz)
            out[f'{tag}_est_{mtag}'] = est
            if mtag == 'all':
                out[f'{tag}_proba'] = m.predict_proba(r)
                out[f'{tag}_labels'] = m.predict_proba_max(r)
    return out


if __name__ == '__main__':
    for name, fn in (('quantizer', quantizer_cases), ('quant', quant_cases), ('gmm', gmm_cases), ('mfa', mfa_cases)):
        d = fn()
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, **d)
        print(name, len(d), 'arrays', os.path.getsize(path), 'bytes')
