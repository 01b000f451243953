"""CPU tests of the host side: precompute vs the reference-generated golden vectors, quantiser tables,
mode dispatch, the C-ABI library's exported symbols (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

from conftest import GMM_MODES, GMM_TAGS, MFA_MODES, MFA_TAGS, ROOT, golden_quantizer_tuple, relerr
import quantized_channel_estimation_b200 as qce
from quantized_channel_estimation_b200 import _lib, build, engine, precompute
from quantized_channel_estimation_b200 import lloyd_max_quantizer as lm
from quantized_channel_estimation_b200 import uniform_quantizer as uq
from torch_ref import reference_combine


def _nb(g, tag):
    nb = float(g[f'{tag}_nbits'])
    return int(nb) if np.isfinite(nb) else np.inf


@pytest.mark.parametrize('tag', GMM_TAGS)
def test_precompute_gmm_vs_reference(golden_gmm, tag):
    g = golden_gmm
    qz = golden_quantizer_tuple(g, tag)
    prep = precompute.prepare(g[f'{tag}_means'], g[f'{tag}_covs'], g[f'{tag}_w'], g[f'{tag}_A'], float(g[f'{tag}_snr']),
                              _nb(g, tag), str(g[f'{tag}_qtype']), qz, device='cpu')
    np.testing.assert_allclose(prep['m_r'].numpy(), g[f'{tag}_mr'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(prep['C_r'].numpy(), g[f'{tag}_Cr'], rtol=1e-13, atol=1e-15)
    for mtag, mode in GMM_MODES.items():
        if f'{tag}_est_{mtag}' not in g:
            continue
        est, lp = reference_combine(prep, g[f'{tag}_r'], mode)
        assert relerr(est.numpy(), g[f'{tag}_est_{mtag}']) < 1e-12
        if mtag == 'all':
            np.testing.assert_allclose(lp.numpy(), g[f'{tag}_wlp'], rtol=1e-12)


@pytest.mark.parametrize('tag', MFA_TAGS)
def test_precompute_mfa_vs_reference(golden_mfa, tag):
    g = golden_mfa
    qz = golden_quantizer_tuple(g, tag)
    prep = precompute.prepare(g[f'{tag}_means'], g[f'{tag}_covs'], g[f'{tag}_amps'], np.eye(8), float(g[f'{tag}_snr']),
                              _nb(g, tag), str(g[f'{tag}_qtype']), qz, device='cpu')
    for mtag, mode in MFA_MODES.items():
        est, lp = reference_combine(prep, g[f'{tag}_r'], mode, top1_exp_argmax=True)
        assert relerr(est.numpy(), g[f'{tag}_est_{mtag}']) < 1e-12
        if mtag == 'all':
            p = np.exp(lp.numpy() - lp.numpy().max(1, keepdims=True))
            np.testing.assert_allclose(p / p.sum(1, keepdims=True), g[f'{tag}_proba'], rtol=1e-10, atol=1e-300)


def test_not_positive_definite_raises_reference_error():
    K, N = 2, 4
    covs = np.stack([np.eye(N, dtype=complex)] * K)
    covs[1] = -covs[1]                                   # C_y = -I + sigma2 I is indefinite at high SNR
    with pytest.raises(ValueError, match='ill-defined empirical covariance'):
        precompute.prepare(np.zeros((K, N), complex), covs, np.ones(K) / K, np.eye(N), 30.0, np.inf, device='cpu')


def test_unknown_quantizer_type():
    with pytest.raises(NotImplementedError):
        qce.get_quantizer([0], 2, 'dither')
    with pytest.raises(NotImplementedError):
        precompute.bussgang_gain(np.ones((1, 2)), 0.0, 2, 'dither', None)


def test_quantizer_tables_match_reference(golden_quantizer):
    g = golden_quantizer
    for nb in (2, 3, 4):
        q = qce.get_quantizer([-10, 0, 10, 20], nb, 'uniform')
        for s in (-10, 0, 10, 20):
            assert np.array_equal(q[s][0], g[f'uni_b{nb}_s{s}_thr'])
            assert np.array_equal(q[s][1], g[f'uni_b{nb}_s{s}_lab'])
    for nb, s in [(2, 0), (2, 10), (3, 0), (3, 10)]:
        thr, lab, rho = lm.load_quantizer(s, nb)[s]
        # closed-form moments vs the reference's scipy.integrate.quad design (its own stop tol is 1e-5)
        np.testing.assert_allclose(thr, g[f'lloyd_b{nb}_s{s}_thr'], atol=1e-9)
        np.testing.assert_allclose(lab, g[f'lloyd_b{nb}_s{s}_lab'], atol=1e-9)
        np.testing.assert_allclose(rho, g[f'lloyd_b{nb}_s{s}_rho'], atol=1e-10)
    assert qce.get_quantizer([0, 5], 1) == {0: (None, None, None), 5: (None, None, None)}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for nb in range(1, 11):
            assert uq.standard_quantization_step(nb) == g[f'step_b{nb}']
            assert uq.standard_distortion_fac(nb) == g[f'rhofac_b{nb}']
            assert uq.get_uniform_quant_step(5, nb) == g[f'qstep_b{nb}_s5']
            assert uq.get_rho_uniform(5, nb) == g[f'rhouni_b{nb}_s5']
            assert lm.get_rho_lloyd(5, nb) == g[f'rholloyd_b{nb}']


def test_bussgang_statistics_match_reference(golden_quantizer):
    g = golden_quantizer
    Cy = g['buss_Cy']
    for nb in (1, 2, 3):
        np.testing.assert_allclose(uq.get_Bussgang_matrix(10, nb, Cy), g[f'buss_uni_b{nb}'], rtol=1e-14)
        qz = qce.get_quantizer([10], nb, 'uniform')[10]
        np.testing.assert_allclose(uq.get_Cr(Cy, nb, 10, qz), g[f'Cr_uni_b{nb}'], rtol=1e-13, atol=1e-16)
    ql = (g['lloyd_b3_s10_thr'], g['lloyd_b3_s10_lab'], None)
    np.testing.assert_allclose(lm.get_Bussgang_matrix(3, Cy, ql), g['buss_lloyd_b3'], rtol=1e-14)
    qu = qce.get_quantizer([10], 2, 'uniform')[10]
    np.testing.assert_allclose(uq.get_quantized_variance(g['qvar_in'], qu), g['qvar_uni_b2'], rtol=1e-14)


def test_mode_dispatch_follows_reference_isinstance_quirk():
    assert engine.parse_mode('all') == (_lib.MODE_ALL, 0, 0.0)
    assert engine.parse_mode(1) == (_lib.MODE_TOP1, 1, 0.0)
    assert engine.parse_mode(3) == (_lib.MODE_TOPN, 3, 0.0)
    assert engine.parse_mode(0.9) == (_lib.MODE_CUMPROB, 0, 0.9)
    assert engine.parse_mode(np.int64(3)) == (_lib.MODE_CUMPROB, 0, 3.0)     # gmm:197 isinstance(x, int)
    with pytest.raises(ValueError):
        engine.parse_mode('some')


def test_data_scale_grid():
    assert precompute.data_scale_for(0, 1, 'uniform') == 1 / np.sqrt(2)
    q = qce.get_quantizer([7], 3, 'uniform')[7]
    m = q[1] / precompute.data_scale_for(7, 3, 'uniform')
    np.testing.assert_allclose(m, np.round(m), atol=1e-12)
    assert set(np.round(m).astype(int)) == {-7, -5, -3, -1, 1, 3, 5, 7}
    assert precompute.data_scale_for(7, 3, 'lloyd') == 0.0
    assert precompute.data_scale_for(7, np.inf, 'uniform') == 0.0


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, 'include', 'qce_b200.h')).read()
    declared = set(re.findall(r'\b(qce_[a-z_0-9]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.qce_abi_version.restype = ctypes.c_int
    assert lib.qce_abi_version() == 1


def test_hot_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    g = qce.Gmm_nbit(n_components=2, covariance_type='full')
    g.set_parameters(np.zeros((2, 4)), np.stack([np.eye(4)] * 2), [0.5, 0.5], zero_mean=True)
    with pytest.raises(RuntimeError, match='no CPU fallback|no B200'):
        g.estimate_from_y(np.ones((3, 4), complex), 0.0, 4, n_summands_or_proba='all')
    with pytest.raises(RuntimeError, match='no CPU fallback|no B200'):
        qce.quant(np.ones((3, 4), complex), 1)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'quantized_channel_estimation_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'(import|from)\s+oracle|oracle[./]|qce_oracle', src), os.path.join(dirpath, f)


def test_load_reference_model_files():
    """Models saved by the reference scripts (joblib .sav of modules.gmm_cplx_bussgang.Gmm_nbit / modules.mofa_cplx_bussgang.Mofa,
    written by tests/golden/make_golden_sav.py with the unmodified classes) load without the reference on the path."""
    import sys
    from quantized_channel_estimation_b200 import Gmm_nbit, Mofa, utils
    assert not any(m == 'modules' or m.startswith('modules.') for m in sys.modules)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_sav_params.npz'))
    gmm = utils.load_reference_model(os.path.join(ROOT, 'tests', 'golden', 'ref_gmm.sav'))
    assert isinstance(gmm, Gmm_nbit)
    np.testing.assert_array_equal(gmm.means_cplx, g['gmm_means'])
    np.testing.assert_array_equal(gmm.covs_cplx, g['gmm_covs'])
    np.testing.assert_array_equal(gmm.gm.weights_, g['w'])
    assert gmm.params['zero_mean'] is False
    mfa = utils.load_reference_model(os.path.join(ROOT, 'tests', 'golden', 'ref_mofa.sav'))
    assert isinstance(mfa, Mofa)
    np.testing.assert_array_equal(mfa.lambdas, g['mfa_lambdas'])
    np.testing.assert_array_equal(mfa.psis, g['mfa_psis'])
    np.testing.assert_array_equal(mfa.covs, g['mfa_covs'])
    np.testing.assert_array_equal(mfa.amps, g['w'])
    assert mfa._covs_are_low_rank


def test_gmm_quant_is_an_inference_alias():
    from quantized_channel_estimation_b200 import Gmm_nbit, Gmm_quant
    g = Gmm_quant(n_components=2, covariance_type='full')
    assert isinstance(g, Gmm_nbit) and Gmm_quant.estimate_from_y is Gmm_nbit.estimate_from_y
    with pytest.raises(NotImplementedError):
        g.fit(np.ones((4, 2), complex), 1, 0.1, None, 'uniform')
    ref = type('Ref', (), {})()
    ref.means_cplx, ref.covs_cplx = np.zeros((2, 3), complex), np.stack([np.eye(3, dtype=complex)] * 2)
    ref.gm, ref.params = type('GM', (), {'weights_': np.array([0.4, 0.6])})(), {'zero_mean': True}
    t = Gmm_quant.from_reference(ref)
    assert isinstance(t, Gmm_quant) and t.params['zero_mean'] is True and t.gm.weights_[1] == 0.6


def test_package_synthetic_generators_match_the_oracle_copies():
    """bench.py / tools/ build their workloads with quantized_channel_estimation_b200.synthetic (product side never imports the
    oracle); the tests use the oracle's own generators: same arrays, seed for seed."""
    from oracle import qce_oracle as orc
    from quantized_channel_estimation_b200 import synthetic
    for a, b in zip(synthetic.random_psd_gmm(3, 8, seed=4, mean_scale=0.2), orc.random_psd_gmm(3, 8, seed=4, mean_scale=0.2)):
        assert np.array_equal(a, b)
    for a, b in zip(synthetic.random_mfa(3, 8, 2, seed=5), orc.random_mfa(3, 8, 2, seed=5)):
        assert np.array_equal(a, b)
    for a, b in zip(synthetic.circulant_gmm(3, 2, 4, seed=6), orc.circulant_gmm(3, 2, 4, seed=6)):
        assert np.array_equal(a, b)
    assert synthetic.circulant_gmm(3, 2, 4, seed=6, dense=False)[1] is None
    means, covs, w = orc.random_psd_gmm(3, 8, seed=4)
    for a, b in zip(synthetic.sample_gmm_channels(means, covs, w, 50, seed=7), orc.sample_gmm_channels(means, covs, w, 50, seed=7)):
        assert np.array_equal(a, b)


def test_toeplitz_helper_matches_scipy():
    import scipy.linalg
    from quantized_channel_estimation_b200 import utils
    rng = np.random.default_rng(0)
    c = rng.standard_normal(6) + 1j * rng.standard_normal(6)
    r = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    assert np.array_equal(utils.toeplitz(c), scipy.linalg.toeplitz(c))
    assert np.array_equal(utils.toeplitz(c, r), scipy.linalg.toeplitz(c, r))


def test_get_pilot_matrix_has_the_reference_signature_and_pilot_types():
    """every reference script calls get_pilot_matrix(n_antennas, n_pilots, n_bits, pilot_type=...) (modules/utils.py:337-367)."""
    import numpy as np
    from quantized_channel_estimation_b200 import utils as u
    x = u.get_pilot_matrix(4, 2, 1, pilot_type='angle_amp', return_vector=True)
    raw = np.array([0.5, 1.0]) * np.exp(1j * np.array([0.0, np.pi / 4]))
    np.testing.assert_allclose(x[:, 0], raw * np.sqrt(2) / np.linalg.norm(raw), atol=1e-15)
    assert abs(np.linalg.norm(x) ** 2 - 2) < 1e-12                      # power constraint
    np.testing.assert_allclose(u.get_pilot_matrix(3, 3, 2, 'angle', True)[:, 0], np.exp(1j * np.array([0, np.pi / 6, np.pi / 3])), atol=1e-15)
    A = u.get_pilot_matrix(3, 2, 1)                                     # positional n_bits, default type
    assert A.shape == (6, 3) and np.allclose(A[:3], x[0, 0] * np.eye(3)) and np.allclose(A[3:], x[1, 0] * np.eye(3))
    assert np.array_equal(u.get_pilot_matrix(5, 1, 1), np.eye(5))       # one pilot: the identity (every script's default)
    assert np.array_equal(u.get_pilot_matrix(2, 3, np.inf, 'angle'), np.kron(np.ones((3, 1)), np.eye(2)))
    assert np.allclose(u.get_pilot_matrix(2, 2, 1, pilots=[1, 1j]), np.kron(np.array([[1], [1j]]), np.eye(2)))
    with pytest.raises(NotImplementedError):
        u.get_pilot_matrix(2, 2, 1, pilot_type='zadoff')
    from oracle import build_ref
    if build_ref.available():                                           # and against the vendored reference itself
        _, _, ru = build_ref.import_reference()
        for pt in ('angle', 'angle_amp', 'ones'):
            for npil in (1, 2, 5):
                np.testing.assert_allclose(u.get_pilot_matrix(4, npil, 2, pilot_type=pt), ru.get_pilot_matrix(4, npil, 2, pilot_type=pt), atol=1e-15)
