"""quantized_channel_estimation_b200 -- B200-native Bussgang-GMM / Bussgang-MFA inference path.

Drop-in for the hot path of benediktfesl/Quantized_Channel_Estimation (``Gmm_nbit`` / ``Mofa``
``estimate_from_y`` and the quantiser helpers) over hand-written sm_100a CUDA kernels reached through
the C ABI in ``include/qce_b200.h``.  Importing the package does not need a GPU; every hot-path call
does, and raises if the library or the device is missing (no CPU fallback).
"""
from . import estimators, lloyd_max_quantizer, uniform_quantizer, utils  # noqa: F401
from .gmm_cplx_bussgang import Gmm_nbit  # noqa: F401
from .gmm_cplx_quant import Gmm_quant  # noqa: F401
from .mofa_cplx_bussgang import Mofa  # noqa: F401
from .utils import get_observation_nbit, get_quantizer, quant  # noqa: F401

__all__ = ['Gmm_nbit', 'Gmm_quant', 'Mofa', 'quant', 'get_observation_nbit', 'get_quantizer', 'utils', 'uniform_quantizer',
           'lloyd_max_quantizer', 'estimators']
