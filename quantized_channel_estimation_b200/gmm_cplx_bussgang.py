"""``Gmm_nbit``: Bussgang-GMM channel estimator for quantised pilots -- the API of the reference's
``modules/gmm_cplx_bussgang.py`` (``fit`` :96-163, ``estimate_from_y`` :166-243) over the CUDA kernels.

Fitted state (``means_cplx [K,N]``, ``covs_cplx [K,N,N]``, ``gm.weights_ [K]``, ``params['zero_mean']``)
has the reference's names so that a model fitted by the reference can be transplanted with
:meth:`Gmm_nbit.from_reference`.  Unlike the reference, ``estimate_from_y`` does not mutate the model:
the per-(SNR, bit width, quantiser, A) parameter blocks are cached as immutable GPU handles.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib, engine, precompute

_SUPPORTED_TYPES = ('full', 'circulant', 'block-circulant', 'toeplitz', 'block-toeplitz')


def _fingerprint(a):
    """Key of a parameter array for the prepared-model cache.  Identity alone would serve stale GPU handles after an in-place edit
    of ``means_cplx`` / ``covs_cplx`` / ``gm.weights_``; a full content hash costs milliseconds per ``estimate_from_y`` call at the
    BASELINE shapes (16 MB of covariances at config 4) while the GPU idles.  So: ``set_parameters`` / ``fit`` make the arrays
    READ-ONLY (an in-place edit raises; assigning a new array changes the identity), and for arrays a caller assigned directly the
    key also carries the bytes of a strided sample of 1024 elements, which follows any edit that touches the array broadly
    (scaling, re-fitting in place)."""
    if a is None:
        return None
    a = np.asarray(a)
    flat = a.reshape(-1)
    sample = flat[::max(1, flat.size // 1024)][:1024]
    return (id(a), a.__array_interface__['data'][0], a.shape, a.dtype.str, hash(sample.tobytes()))


def _frozen(a, dtype):
    """A private read-only copy: the prepared-model cache may then key on identity (see _fingerprint)."""
    a = np.array(a, dtype=dtype)
    a.setflags(write=False)
    return a


def _table_key(quantizer):
    if quantizer is None or quantizer[0] is None:
        return None
    return (np.asarray(quantizer[0], dtype=np.float64).tobytes(), np.asarray(quantizer[1], dtype=np.float64).tobytes())


class _PreparedCache:
    """LRU of DenseModel handles keyed by everything _prepare_for_prediction depends on."""

    def __init__(self, capacity=32):
        self.capacity = capacity
        self.items = {}

    def get(self, key, make):
        if key in self.items:
            self.items[key] = self.items.pop(key)
            return self.items[key]
        val = make()
        self.items[key] = val
        while len(self.items) > self.capacity:
            self.items.pop(next(iter(self.items)))
        return val

    def clear(self):
        self.items.clear()


class Gmm_nbit:
    def __init__(self, *gmm_args, **gmm_kwargs):
        names = ('n_components', 'covariance_type', 'tol', 'reg_covar', 'max_iter', 'n_init', 'init_params')
        kw = dict(n_components=1, covariance_type='full', tol=1e-3, reg_covar=1e-6, max_iter=100, n_init=1,
                  init_params='kmeans', random_state=None, verbose=0)
        kw.update(dict(zip(names, gmm_args)))
        kw.update(gmm_kwargs)
        # attribute container with sklearn.mixture.GaussianMixture's names (the reference keeps one, gmm:87)
        self.gm = SimpleNamespace(weights_=None, means_=None, covariances_=None, precisions_cholesky_=None,
                                  converged_=False, n_iter_=0, lower_bound_=-np.inf, **kw)
        self.means_cplx = None
        self.covs_cplx = None
        self.fft_covs = None
        self.fft_means = None
        self.chol = None
        self.params = dict()
        self.F2 = None
        self.precision = 'auto'            # 'auto' | 'tc' | 'fp64' (arithmetic of the estimate kernel)
        self.blocks = None                 # (n1, n2) when every covariance is F^H diag(c) F with F = F_n1 (x) F_n2
        self.use_structure = True          # use the DFT-domain kernel when blocks is set, A = I and the means vanish
        self._cache = _PreparedCache()
        self._last = None                  # handle prepared by the last estimate_from_y (for predict_proba_cplx(X))

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_reference(cls, obj):
        """Transplant a fitted reference ``Gmm_nbit`` (or anything with the same attributes)."""
        new = cls(n_components=int(np.asarray(obj.means_cplx).shape[0]), covariance_type='full')
        new.set_parameters(obj.means_cplx, obj.covs_cplx, obj.gm.weights_, zero_mean=obj.params.get('zero_mean', False))
        return new

    def set_parameters(self, means, covs, weights, zero_mean=False, detect_structure=True):
        """Install fitted parameters.  ``detect_structure``: test whether the covariances are (block-)circulant
        (the reference fits those types in the DFT domain and then densifies them, gmm:104-136) and, if so, remember
        the DFT-domain eigenvalues so that ``estimate_from_y`` can use the structured kernel."""
        self.means_cplx = _frozen(means, complex)
        self.covs_cplx = _frozen(covs, complex)
        self.gm.weights_ = _frozen(weights, float)
        self.blocks, self.fft_covs = (None, None)
        if detect_structure and self.covs_cplx.shape[-1] <= 1024:
            self.blocks, self.fft_covs = precompute.detect_blocks(self.covs_cplx)
        self.gm.n_components = self.means_cplx.shape[0]
        self.gm.covariance_type = 'full'       # every type is dense after fit (gmm:110-153)
        self.params['zero_mean'] = bool(zero_mean)
        self._cache.clear()
        return self

    def fit(self, h, blocks=None, zero_mean=False):
        """Fit the complex GMM with EM (reference :96-163); see ``em.py``.  'full', 'circulant', 'block-circulant'
        (``blocks=(n1, n2)``) and 'toeplitz' / 'block-toeplitz' (inverse EM on the oversampled DFT grid, reference :792-826)."""
        if self.gm.covariance_type not in _SUPPORTED_TYPES:
            raise NotImplementedError(f'Fitting for covariance_type = {self.gm.covariance_type} is not implemented.')
        from . import em
        em.fit_gmm(self, h, blocks=blocks, zero_mean=zero_mean)
        self._cache.clear()
        return self

    # ------------------------------------------------------------------ inference
    def set_circulant_parameters(self, c, weights, blocks):
        """Zero-mean (block-)circulant mixture from its DFT-domain eigenvalues ``c [K, N]``: ``C_k = F^H diag(c_k) F``,
        ``F = F_n1 (x) F_n2``, ``blocks = (n1, n2)`` (``(1, N)``: circulant).  The dense covariances are materialised too
        (reference behaviour) unless N > 1024."""
        c = np.asarray(c, dtype=float)
        K, N = c.shape
        assert blocks[0] * blocks[1] == N
        self.means_cplx = np.zeros((K, N), dtype=complex)
        if N <= 1024:
            F = precompute.dft_matrix(*blocks)
            self.covs_cplx = np.einsum('ji,kj,jl->kil', F.conj(), c, F)
        else:
            self.covs_cplx = None
        self.gm.weights_ = np.array(weights, dtype=float)
        self.gm.n_components = K
        self.gm.covariance_type = 'full'
        self.params['zero_mean'] = True
        self.blocks, self.fft_covs = tuple(blocks), c.copy()
        self._cache.clear()
        return self

    def _structured(self, A):
        if not (self.use_structure and self.blocks is not None and self.fft_covs is not None):
            return False
        A = np.asarray(A)
        K, N = self.fft_covs.shape
        if max(self.blocks) > 256 or K > 4096:       # limits of qce_circ_model_create; larger models take the dense path
            return False
        return A.shape == (N, N) and np.array_equal(A, np.eye(N)) and not np.any(self.means_cplx)

    def _prepared(self, A, snr_dB, n_bits, quantizer_type, quantizer):
        if self._structured(A):
            nb = 'inf' if (n_bits == 'inf' or n_bits == np.inf) else int(n_bits)
            tables = _table_key(quantizer) if (nb != 1 and nb != 'inf' and quantizer_type == 'lloyd') else None
            key = ('circ', float(snr_dB), nb, quantizer_type if nb not in (1, 'inf') else None, tables, self.blocks,
                   _fingerprint(self.fft_covs), _fingerprint(self.gm.weights_))

            def make_circ():
                prep = precompute.prepare_circulant(self.fft_covs, self.gm.weights_, self.blocks, snr_dB,
                                                    np.inf if nb == 'inf' else nb, quantizer_type, quantizer)
                return engine.CircModel(prep, flags=0)
            try:
                return self._cache.get(key, make_circ)
            except _lib.QceError as e:      # a shape the structured kernels refuse: the dense path handles it (reference behaviour)
                if e.status not in (_lib.ERR_INVALID, _lib.ERR_UNSUPPORTED) or self.covs_cplx is None:
                    raise
        if self.means_cplx is None or self.covs_cplx is None or self.gm.weights_ is None:
            raise RuntimeError('Gmm_nbit: model is not fitted (means_cplx / covs_cplx / gm.weights_ missing)')
        if self.gm.covariance_type != 'full':
            raise NotImplementedError(f'Estimation for covariance_type = {self.gm.covariance_type} is not implemented.')
        A = np.asarray(A)
        nb = 'inf' if (n_bits == 'inf' or n_bits == np.inf) else int(n_bits)
        tables = _table_key(quantizer) if (nb != 1 and nb != 'inf' and quantizer_type == 'lloyd') else None
        pad = self.precision != 'fp64'        # shapes the tensor-core kernels do not cover are zero-padded to one they do
        key = (float(snr_dB), nb, quantizer_type if nb not in (1, 'inf') else None, tables, A.shape, A.tobytes(),
               _fingerprint(self.means_cplx), _fingerprint(self.covs_cplx), _fingerprint(self.gm.weights_), pad)

        def make():
            prep = precompute.prepare(self.means_cplx, self.covs_cplx, self.gm.weights_, A, snr_dB,
                                      np.inf if nb == 'inf' else nb, quantizer_type, quantizer)
            return engine.DenseModel(prep, flags=0, pad=pad)
        return self._cache.get(key, make)

    def estimate_from_y(self, y, snr_dB, n_antennas, A=None, n_summands_or_proba=1, n_bits=1,
                        quantizer_type='uniform', quantizer=None):
        """Channel estimates ``[B, A.shape[-1]]`` complex128 from quantised pilots ``y [B, n_obs]``
        (reference :166-243; arguments have the reference's meaning).  ``y`` may be a torch CUDA tensor
        (result: CUDA tensor) or a numpy array / CPU tensor (result: numpy array / CPU tensor; the
        host<->device copies are chunked and overlapped inside the library)."""
        if A is None:
            A = np.eye(n_antennas, dtype=complex)
        model = self._last = self._prepared(A, snr_dB, n_bits, quantizer_type, quantizer)
        if isinstance(y, torch.Tensor):
            if y.is_cuda:
                return model.estimate(y, n_summands_or_proba, self.precision)
            return torch.from_numpy(model.estimate_host(y.numpy(), n_summands_or_proba, self.precision))
        return model.estimate_host(y, n_summands_or_proba, self.precision)

    def estimate_from_codes(self, codes, snr_dB, n_antennas, A=None, n_summands_or_proba=1, n_bits=1, quantizer_type='uniform',
                            quantizer=None, out_dtype=np.complex64):
        """``estimate_from_y`` for pilots given as quantiser LEVEL CODES on the host (uint8 ``[B, n_obs, 2]``: per real dimension the
        index of the level; 1 bit: 0 negative / 1 zero / 2 positive) with complex64 estimates back by default: 2 B per pilot entry in
        and 8 B out cross PCIe instead of 16 B + 16 B.  Same kernels, same results (narrowed to ``out_dtype``)."""
        if A is None:
            A = np.eye(n_antennas, dtype=complex)
        model = self._last = self._prepared(A, snr_dB, n_bits, quantizer_type, quantizer)
        if not isinstance(model, engine.DenseModel):
            raise NotImplementedError('estimate_from_codes: dense models only (set use_structure = False)')
        q = engine.Quantizer.get(1) if int(n_bits) == 1 else engine.Quantizer.get(int(n_bits), quantizer[0], quantizer[1])
        return model.estimate_host_codes(codes, q, n_summands_or_proba, self.precision, out_dtype)

    def weighted_log_prob(self, y, snr_dB=None, A=None, n_bits=1, quantizer_type='uniform', quantizer=None):
        """``_estimate_weighted_log_prob`` of the prepared mixture (reference :369-386): ``[B, K]`` float64.  With ``snr_dB=None``
        the setting prepared by the most recent ``estimate_from_y`` is used -- the reference's calling convention, whose
        ``predict_proba_cplx(X)`` reads the state ``_prepare_for_prediction`` left in ``self.gm``."""
        if snr_dB is None:
            if self._last is None:
                # no observation setting prepared yet: the trained mixture itself (the state fit() leaves in the reference's self.gm)
                if self.means_cplx is None or self.covs_cplx is None:
                    raise RuntimeError('Gmm_nbit: the model is not fitted (means_cplx / covs_cplx missing)')
                self._last = self._prepared(np.eye(self.means_cplx.shape[1], dtype=complex), np.inf, np.inf, 'uniform', None)
            model = self._last
        else:
            if A is None:
                A = np.eye(self.means_cplx.shape[1], dtype=complex)
            model = self._last = self._prepared(A, snr_dB, n_bits, quantizer_type, quantizer)
        yt = y if isinstance(y, torch.Tensor) and y.is_cuda else torch.as_tensor(np.asarray(y)).cuda()
        # the whitening launch of the estimate path (self.precision: tensor cores when the shape is covered), log-probabilities only
        lp = model.log_prob(yt, self.precision)
        return lp if isinstance(y, torch.Tensor) and y.is_cuda else lp.cpu().numpy()

    _estimate_weighted_log_prob = weighted_log_prob

    def predict_proba_cplx(self, y, snr_dB=None, A=None, n_bits=1, quantizer_type='uniform', quantizer=None):
        """Responsibilities ``p(k | r)`` of the prepared mixture (reference :351-367)."""
        lp = self.weighted_log_prob(y, snr_dB, A, n_bits, quantizer_type, quantizer)
        if isinstance(lp, torch.Tensor):
            return torch.softmax(lp, dim=1)
        m = lp.max(axis=1, keepdims=True)
        e = np.exp(lp - m)
        return e / e.sum(axis=1, keepdims=True)

    def _predict_cplx(self, y, snr_dB=None, A=None, n_bits=1, quantizer_type='uniform', quantizer=None):
        """Hard labels: argmax of the weighted log-probabilities (reference :335-349)."""
        lp = self.weighted_log_prob(y, snr_dB, A, n_bits, quantizer_type, quantizer)
        return lp.argmax(dim=1) if isinstance(lp, torch.Tensor) else lp.argmax(axis=1)
