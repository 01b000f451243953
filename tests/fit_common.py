"""Seeded data generators and the likelihood evaluator shared by tests/golden/make_golden_fit.py and tests/test_fit_cpu.py."""
import numpy as np


def crandn(rng, *shape):
    return np.sqrt(0.5) * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def avg_loglik(h, weights, means, covs):
    """Mean log-likelihood of ``h [B,N]`` under a complex Gaussian mixture (numpy, float64)."""
    B, N = h.shape
    lp = np.empty((B, len(weights)))
    for k in range(len(weights)):
        d = h - means[k]
        _, logdet = np.linalg.slogdet(covs[k])
        q = np.real(np.sum(d.conj() * np.linalg.solve(covs[k], d.T).T, axis=1))
        lp[:, k] = np.log(weights[k]) - N * np.log(np.pi) - logdet - q
    m = lp.max(1)
    return float(np.mean(m + np.log(np.exp(lp - m[:, None]).sum(1))))


def make_data(tag, B=4000):
    """Samples of a known 3-component mixture; returns ``(h, (weights, means, covs))``."""
    rng = np.random.default_rng({'full_zm': 1, 'full_mean': 2, 'circ': 3, 'bccb': 4, 'mfa': 5, 'toep': 6, 'btoep': 7}[tag])
    K = 3
    w = np.array([0.5, 0.3, 0.2])
    if tag in ('full_zm', 'full_mean'):
        N = 4
        covs = []
        for k in range(K):
            X = crandn(rng, N, 2 * N)
            C = X @ X.conj().T / (2 * N) * (0.3 + k)
            covs.append(C)
        means = np.zeros((K, N), complex) if tag == 'full_zm' else 1.5 * crandn(rng, K, N)
    elif tag in ('toep', 'btoep'):
        N = 8

        def toep(n):            # positive definite Toeplitz matrix from a few spectral lines plus a white floor
            f, p = rng.random(3), rng.random(3) + 0.2
            t = (p[None, :] * np.exp(2j * np.pi * f[None, :] * np.arange(n)[:, None])).sum(1)
            t[0] += 0.2
            idx = np.arange(n)[None, :] - np.arange(n)[:, None]
            return np.where(idx >= 0, t[np.abs(idx)], np.conj(t[np.abs(idx)]))
        covs = [(0.3 + k) * (toep(8) if tag == 'toep' else np.kron(toep(2), toep(4))) for k in range(K)]
        means = np.zeros((K, N), complex)
    elif tag in ('circ', 'bccb'):
        N = 8
        n1, n2 = (1, 8) if tag == 'circ' else (2, 4)
        F = np.kron(np.fft.fft(np.eye(n1)) / np.sqrt(n1), np.fft.fft(np.eye(n2)) / np.sqrt(n2))
        covs = [F.conj().T @ np.diag((rng.random(N) ** 2 + 0.05) * (0.3 + 2 * k)) @ F for k in range(K)]
        means = np.zeros((K, N), complex)
    else:
        N, M = 8, 2
        covs = []
        for k in range(K):
            lam = crandn(rng, N, M) * (0.5 + k)
            covs.append(lam @ lam.conj().T + np.diag(0.05 + 0.1 * rng.random(N)))
        means = 2.0 * crandn(rng, K, N)
    covs = np.stack(covs)
    lab = rng.choice(K, size=B, p=w)
    h = np.empty((B, covs.shape[-1]), complex)
    for b in range(B):
        L = np.linalg.cholesky(covs[lab[b]])
        h[b] = means[lab[b]] + L @ crandn(rng, covs.shape[-1])
    return h, (w, means, covs)
