"""Pin ``oracle/qce_oracle.py`` against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import warnings

import numpy as np
import pytest

from conftest import BASELINE_TAGS, GMM_MODES, GMM_TAGS, GMM_TC_TAGS, MFA_MODES, MFA_TAGS, baseline_case, golden_quantizer_tuple, relerr
from oracle import qce_oracle as orc


def test_known_answer_constants():
    # constants probed from the reference
    assert (1 / np.sqrt(2)).hex() == float.fromhex('0x1.6a09e667f3bccp-1').hex()
    assert orc.get_uniform_quant_step(10, 2) == 0.7384308833601152
    q = orc.get_quantizer([10], 2, 'uniform')[10]
    np.testing.assert_allclose(q[0], [-0.73843088, 0, 0.73843088], atol=1e-8)
    np.testing.assert_allclose(q[1], [-1.10764633, -0.36921544, 0.36921544, 1.10764633], atol=1e-8)


def test_quantizer_tables(golden_quantizer):
    g = golden_quantizer
    for nb in (2, 3, 4):
        q = orc.get_quantizer([-10, 0, 10, 20], nb, 'uniform')
        for s in (-10, 0, 10, 20):
            assert np.array_equal(q[s][0], g[f'uni_b{nb}_s{s}_thr'])
            assert np.array_equal(q[s][1], g[f'uni_b{nb}_s{s}_lab'])
            assert q[s][2] is None
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for nb in range(1, 11):
            assert orc.standard_quantization_step(nb) == g[f'step_b{nb}']
            assert orc.standard_distortion_fac(nb) == g[f'rhofac_b{nb}']
            assert orc.get_uniform_quant_step(5, nb) == g[f'qstep_b{nb}_s5']
            assert orc.get_rho_uniform(5, nb) == g[f'rhouni_b{nb}_s5']
            assert orc.get_rho_lloyd(5, nb) == g[f'rholloyd_b{nb}']


@pytest.mark.parametrize('nb,snr', [(2, 0), (2, 10), (3, 0), (3, 10)])
def test_lloyd_design(golden_quantizer, nb, snr):
    g = golden_quantizer
    thr, lab, rho = orc.load_quantizer(snr, nb)[snr]
    assert np.array_equal(thr, g[f'lloyd_b{nb}_s{snr}_thr'])
    assert np.array_equal(lab, g[f'lloyd_b{nb}_s{snr}_lab'])
    assert rho == g[f'lloyd_b{nb}_s{snr}_rho']


def test_bussgang_and_cr(golden_quantizer):
    g = golden_quantizer
    Cy = g['buss_Cy']
    for nb in (1, 2, 3):
        np.testing.assert_allclose(orc.uniform_get_bussgang_matrix(10, nb, Cy), g[f'buss_uni_b{nb}'], rtol=1e-15, atol=0)
        qz = orc.get_quantizer([10], nb, 'uniform')[10]
        np.testing.assert_allclose(orc.get_Cr(Cy, nb, 10, qz), g[f'Cr_uni_b{nb}'], rtol=1e-14, atol=1e-16)
    ql = (g['lloyd_b3_s10_thr'], g['lloyd_b3_s10_lab'], None)
    np.testing.assert_allclose(orc.lloyd_get_bussgang_matrix(3, Cy, ql), g['buss_lloyd_b3'], rtol=1e-15, atol=0)
    qu = orc.get_quantizer([10], 2, 'uniform')[10]
    np.testing.assert_allclose(orc.get_quantized_variance(g['qvar_in'], qu), g['qvar_uni_b2'], rtol=1e-15)


def _bits_equal(a, b):
    a = np.ascontiguousarray(a).view(np.uint64)
    b = np.ascontiguousarray(b).view(np.uint64)
    return np.array_equal(a, b)


def test_quant_bit_exact(golden_quant):
    g = golden_quant
    assert _bits_equal(orc.quant(g['y'], 1), g['q1'])
    qu = orc.get_quantizer([10], 2, 'uniform')[10]
    assert _bits_equal(orc.quant(g['yu'], 2, qu[0], qu[1]), g['q2u'])
    assert _bits_equal(orc.quant(g['y'], 3, g['q3l_thr'], g['q3l_lab']), g['q3l'])
    # integer codes reproduce the labels
    codes = orc.quant_codes(g['y'], 3, g['q3l_thr'])
    assert _bits_equal(g['q3l_lab'][codes[..., 0]] + 1j * g['q3l_lab'][codes[..., 1]], g['q3l'])
    c1 = orc.quant_codes(g['y'], 1)
    sgn = np.array([-1.0, 0.0, 1.0, np.nan])
    assert _bits_equal(1 / np.sqrt(2) * (sgn[c1[..., 0]] + 1j * sgn[c1[..., 1]]), g['q1'])


def test_observation_bit_exact(golden_quant):
    g = golden_quant
    for snr in (-5, 10):
        assert _bits_equal(orc.get_observation_nbit(g['obs_h'], snr, g['obs_noise'], None, 1), g[f'obs1_s{snr}'])
        qs = orc.get_quantizer([snr], 2, 'uniform')[snr]
        assert _bits_equal(orc.get_observation_nbit(g['obs_h'], snr, g['obs_noise'], None, 2, qs[0], qs[1]),
                           g[f'obs2u_s{snr}'])
        assert _bits_equal(orc.get_observation_nbit(g['obs_h'], snr, g['obs_noise'], None, np.inf),
                           g[f'obsinf_s{snr}'])


@pytest.mark.parametrize('tag', GMM_TAGS)
def test_gmm_estimate(golden_gmm, tag):
    g = golden_gmm
    nb = float(g[f'{tag}_nbits'])
    nb = int(nb) if np.isfinite(nb) else np.inf
    qz = golden_quantizer_tuple(g, tag)
    snr = float(g[f'{tag}_snr'])
    # the quantised pilots themselves (observation synthesis incl. A != I)
    r = orc.get_observation_nbit(g[f'{tag}_h'], snr, g[f'{tag}_noise'], g[f'{tag}_A'], nb, qz[0], qz[1])
    assert _bits_equal(r, g[f'{tag}_r'])
    for mtag, mode in GMM_MODES.items():
        key = f'{tag}_est_{mtag}'
        if key not in g:
            continue
        est, aux = orc.gmm_estimate_from_y(g[f'{tag}_means'], g[f'{tag}_covs'], g[f'{tag}_w'], g[f'{tag}_r'], snr,
                                           A=g[f'{tag}_A'], n_summands_or_proba=mode, n_bits=nb,
                                           quantizer_type=str(g[f'{tag}_qtype']), quantizer=qz, return_aux=True)
        assert relerr(est, g[key]) < 1e-12, (tag, mtag)
        if mtag == 'all':
            np.testing.assert_allclose(aux['proba'], g[f'{tag}_proba'], rtol=1e-10, atol=1e-300)
            np.testing.assert_allclose(aux['wlp'], g[f'{tag}_wlp'], rtol=1e-12)
            np.testing.assert_allclose(aux['prep']['m_r'], g[f'{tag}_mr'], rtol=1e-14, atol=1e-16)
            np.testing.assert_allclose(aux['prep']['C_r'], g[f'{tag}_Cr'], rtol=1e-14, atol=1e-16)


@pytest.mark.parametrize('tag', GMM_TC_TAGS)
def test_gmm_estimate_tc_shapes(golden_gmm_tc, tag):
    """The oracle against the reference outputs at the tensor-core shapes (N = 16 / 32)."""
    g = golden_gmm_tc
    nb = float(g[f'{tag}_nbits'])
    nb = int(nb) if np.isfinite(nb) else np.inf
    qz = golden_quantizer_tuple(g, tag)
    for mtag, mode in GMM_MODES.items():
        est, aux = orc.gmm_estimate_from_y(g[f'{tag}_means'], g[f'{tag}_covs'], g[f'{tag}_w'], g[f'{tag}_r'], float(g[f'{tag}_snr']),
                                           A=g[f'{tag}_A'], n_summands_or_proba=mode, n_bits=nb,
                                           quantizer_type=str(g[f'{tag}_qtype']), quantizer=qz, return_aux=True)
        assert relerr(est, g[f'{tag}_est_{mtag}']) < 1e-11, (tag, mtag)
        if mtag == 'all':
            np.testing.assert_allclose(aux['wlp'], g[f'{tag}_wlp'], rtol=1e-11)


@pytest.mark.parametrize('tag', MFA_TAGS)
def test_mfa_estimate(golden_mfa, tag):
    g = golden_mfa
    nb = int(g[f'{tag}_nbits'])
    qz = golden_quantizer_tuple(g, tag)
    snr = float(g[f'{tag}_snr'])
    np.testing.assert_allclose(orc.mofa_covs(g[f'{tag}_lambdas'], g[f'{tag}_psis']), g[f'{tag}_covs'], rtol=1e-14)
    for mtag, mode in MFA_MODES.items():
        est, aux = orc.mofa_estimate_from_y(g[f'{tag}_means'], g[f'{tag}_covs'], g[f'{tag}_amps'], g[f'{tag}_r'], snr,
                                            n_summands_or_proba=mode, n_bits=nb,
                                            quantizer_type=str(g[f'{tag}_qtype']), quantizer=qz, return_aux=True)
        assert relerr(est, g[f'{tag}_est_{mtag}']) < 1e-12, (tag, mtag)
        if mtag == 'all':
            np.testing.assert_allclose(aux['proba'], g[f'{tag}_proba'], rtol=1e-10, atol=1e-300)
            assert np.array_equal(np.exp(aux['logrs']).argmax(axis=0), g[f'{tag}_labels'])


@pytest.mark.parametrize('tag', BASELINE_TAGS)
def test_baselines(golden_baselines, tag):
    """Global / genie Bussgang-LMMSE and Bussgang-LS (estimators/blmmse.py, estimators/LS.py) against the reference's outputs."""
    g = golden_baselines
    r, A, nb, qt, qz = baseline_case(g, tag)
    snr = float(g['snr'])
    assert relerr(orc.blmmse_estimate_global(r, g['C_glob'], snr, A, nb, qt, qz), g[tag + '_blmmse_global']) < 1e-12
    assert relerr(orc.ls_estimate_global(r, g['C_glob'], snr, A, nb, qt, qz), g[tag + '_ls_global']) < 1e-12
    assert relerr(orc.blmmse_estimate_genie(r, g['t'], snr, A, nb, qt, qz), g[tag + '_blmmse_genie']) < 1e-11
    if tag + '_ls_genie' in g:
        assert relerr(orc.ls_estimate_genie(r, g['t'], snr, A, nb, qt, qz), g[tag + '_ls_genie']) < 1e-12


def test_oracle_vs_vendored_reference():
    """Second pin of the oracle: the UNMODIFIED reference modules vendored by oracle/build_ref.py (where they exist), run here on a
    fresh seeded case that is in no fixture -- all four combination modes, 2-bit uniform pilots, non-zero means."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip('oracle/_ref not built (run python oracle/build_ref.py where /root/reference exists)')
    Gmm_nbit, _, ut = build_ref.import_reference()
    K, N, B, snr, nb = 5, 12, 40, 7, 2
    means, covs, w = orc.random_psd_gmm(K, N, seed=31, mean_scale=0.2)
    h, noise, _ = orc.sample_gmm_channels(means, covs, w, B, seed=32)
    qz = ut.get_quantizer_gauss([snr], nb, 'uniform')[snr]
    r = orc.get_observation_nbit(h, snr, noise, None, nb, qz[0], qz[1])
    assert _bits_equal(r, ut.quant(h + 10 ** (-snr / 20) * noise, nb, qz[0], qz[1]))
    for mode in ('all', 1, 2, 0.8):
        g = Gmm_nbit(n_components=K, covariance_type='full')
        g.params['zero_mean'] = False
        g.means_cplx, g.covs_cplx = means.copy(), covs.copy()
        g.gm.weights_ = w.copy()
        ref = g.estimate_from_y(r, snr, N, None, mode, nb, 'uniform', qz)
        est = orc.gmm_estimate_from_y(means, covs, w, r, snr, n_summands_or_proba=mode, n_bits=nb, quantizer_type='uniform', quantizer=qz)
        assert relerr(est, ref) < 1e-12, mode
