"""Seeded synthetic parameter / data generators of SURVEY.md section 8d (host numpy): random-PSD GMMs, random MFA models,
(block-)circulant GMMs and mixture samples.  Used by ``bench.py`` and ``tools/`` to build workloads of the BASELINE shapes
(no datasets or checkpoints exist offline); they produce the same arrays, seed for seed, as the generators the test oracle
keeps for itself (checked by tests/test_host_cpu.py)."""
import numpy as np


def crandn(*shape, rng):
    """Circularly-symmetric complex standard normal draw from ``rng``."""
    return np.sqrt(0.5) * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def random_psd_gmm(K, N, seed=0, mean_scale=0.0):
    """P-rand: ``C = X X^H / (2N)`` with ``X ~ CN(0,1)^{N x 2N}``, scaled to ``tr C = N``; ``w = U(0,1)^K / sum``."""
    rng = np.random.default_rng(seed)
    covs = np.empty((K, N, N), dtype=complex)
    for k in range(K):
        X = crandn(N, 2 * N, rng=rng)
        C = X @ X.conj().T / (2 * N)
        C = C * (N / np.real(np.trace(C)))
        covs[k] = 0.5 * (C + C.conj().T)
    w = rng.random(K)
    w = w / w.sum()
    means = mean_scale * crandn(K, N, rng=rng) if mean_scale else np.zeros((K, N), dtype=complex)
    return means, covs, w


def random_mfa(K, N, M, seed=0, mean_scale=0.0):
    """P-mfa: ``Lambda ~ CN(0, 1/M)^{K x N x M}``, ``psi ~ 0.02 + 0.1 U(0,1)``, amplitudes like the GMM weights."""
    rng = np.random.default_rng(seed)
    lambdas = crandn(K, N, M, rng=rng) / np.sqrt(M)
    psis = 0.02 + 0.1 * rng.random((K, N))
    amps = rng.random(K)
    amps = amps / amps.sum()
    means = mean_scale * crandn(K, N, rng=rng) if mean_scale else np.zeros((K, N), dtype=complex)
    return means, lambdas, psis, amps


def circulant_gmm(K, n1, n2, seed=0, dense=True):
    """P-circ / BCCB: ``c_k = U(0,1)^N ** 3 + 1e-3`` mean-normalised, ``C_k = F^H diag(c_k) F``, ``F = F_n1 (x) F_n2`` unitary.
    Returns ``(c, covs or None, w, F)``."""
    rng = np.random.default_rng(seed)
    N = n1 * n2
    c = rng.random((K, N)) ** 3 + 1e-3
    c = c / c.mean(axis=1, keepdims=True)
    F = np.kron(np.fft.fft(np.eye(n1)) / np.sqrt(n1), np.fft.fft(np.eye(n2)) / np.sqrt(n2))
    covs = np.einsum('ji,kj,jl->kil', F.conj(), c, F) if dense else None
    w = rng.random(K)
    w = w / w.sum()
    return c, covs, w, F


def sample_gmm_channels(means, covs, weights, B, seed=1):
    """``k_b ~ Cat(w)``, ``h_b = mu_k + C_k^{1/2} g_b``; returns ``(h [B,N], noise [B,N], labels)`` (complex128)."""
    rng = np.random.default_rng(seed)
    K, N = means.shape
    lab = rng.choice(K, size=B, p=weights)
    h = np.empty((B, N), dtype=complex)
    for k in np.unique(lab):
        idx = np.nonzero(lab == k)[0]
        L = np.linalg.cholesky(covs[k] + 1e-12 * np.eye(N))
        h[idx] = means[k] + crandn(idx.size, N, rng=rng) @ L.T
    noise = crandn(B, N, rng=rng)
    return h, noise, lab
