"""``Gmm_quant`` -- the inference surface of the reference's ``modules/gmm_cplx_quant.py``.

The reference class differs from ``Gmm_nbit`` in how it is TRAINED (``fit(h, n_bits, sigma2, quantizer, quant_type, ...)``,
gmm_cplx_quant.py:103-188: EM on quantised data with covariance recovery).  Its inference path -- ``estimate_from_y``
(:190-268), ``_prepare_for_prediction`` (:270-348), ``predict_proba_cplx`` / ``_predict_cplx`` (:354-386) -- is line for line the
one of ``Gmm_nbit`` (without the infinite-resolution branch), so a model fitted by the reference and transplanted with
``Gmm_quant.from_reference`` / ``utils.load_reference_model`` runs on the same CUDA kernels.  Training from quantised data is out
of scope of this package and raises.
"""
from .gmm_cplx_bussgang import Gmm_nbit


class Gmm_quant(Gmm_nbit):
    def fit(self, h, n_bits=None, sigma2=None, quantizer=None, quant_type=None, blocks=None, zero_mean=False):
        raise NotImplementedError('Gmm_quant.fit (EM on quantised data with covariance recovery, gmm_cplx_quant.py:103-188) is not '
                                  'part of this package; fit with the reference and transplant the model with '
                                  'Gmm_quant.from_reference(obj) or utils.load_reference_model(path)')
