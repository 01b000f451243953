#!/usr/bin/env python
"""End-to-end rate of the reference-facing call (numpy arrays in, numpy array out) for pageable and page-locked host buffers."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantized_channel_estimation_b200 as qce
from quantized_channel_estimation_b200 import synthetic


def rate(fn, B, N, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    return dict(est_per_s=B / dt, gbytes_per_s_each_way=B * N * 16 / dt / 1e9)


def main():
    snr = 10
    # config 2: dense GMM, 1 bit, N = 64, K = 64
    K, N, B = 64, 64, 1 << 19
    means, covs, w = synthetic.random_psd_gmm(K, N, seed=0)
    m = qce.Gmm_nbit(n_components=K).set_parameters(means, covs, w, detect_structure=False)
    g = torch.Generator(device='cuda').manual_seed(0)
    bits = torch.randint(0, 2, (B, N, 2), generator=g, device='cuda', dtype=torch.int8)
    r_dev = torch.view_as_complex(((bits.double() * 2 - 1) / np.sqrt(2)).contiguous())
    r = r_dev.cpu().numpy()
    print(json.dumps(dict(case='C2 dense, pageable numpy', **rate(lambda: m.estimate_from_y(r, snr, N, n_summands_or_proba='all'), B, N))), flush=True)
    rp = torch.empty((B, N), dtype=torch.complex128).pin_memory()
    rp.copy_(r_dev)
    print(json.dumps(dict(case='C2 dense, pinned input (output array still pageable)', **rate(lambda: m.estimate_from_y(rp.numpy(), snr, N, n_summands_or_proba='all'), B, N))), flush=True)
    # config 3: block-circulant, 3-bit Lloyd-Max, N = 256, K = 128
    K, N, B = 128, 256, 1 << 17
    c, _, w, _ = synthetic.circulant_gmm(K, 16, 16, seed=0)
    qz = qce.get_quantizer([snr], 3, 'lloyd')[snr]
    mc = qce.Gmm_nbit(n_components=K, covariance_type='block-circulant')
    mc.set_circulant_parameters(c, w, (16, 16))
    y = torch.view_as_complex(torch.randn((B, N, 2), generator=g, device='cuda', dtype=torch.float64)) * 0.8
    r3 = qce.quant(y, 3, qz[0], qz[1]).cpu().numpy()
    print(json.dumps(dict(case='C3 block-circulant, pageable numpy', **rate(
        lambda: mc.estimate_from_y(r3, snr, N, n_summands_or_proba='all', n_bits=3, quantizer_type='lloyd', quantizer=qz), B, N))), flush=True)


if __name__ == '__main__':
    main()
