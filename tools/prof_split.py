#!/usr/bin/env python
"""GMM 'full', 1 bit, N = 128, K = 64 (the tensor-core split path: whitening-only launch -> selection -> two LMMSE row-block
launches) alone: timing loop for ncu / quick A-B runs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantized_channel_estimation_b200 as qce
from quantized_channel_estimation_b200 import synthetic
from bench_configs import pilots, timeit


def main():
    snr, K, N = 10, int(os.environ.get('K', 64)), int(os.environ.get('N', 128))
    B = 1 << int(os.environ.get('LOG2B', 19))
    means, covs, w = synthetic.random_psd_gmm(K, N, seed=0)
    m = qce.Gmm_nbit(n_components=K).set_parameters(means, covs, w, detect_structure=False)
    m.precision = 'tc'
    r = pilots(B, N, 1, (None, None))
    for mode in ('all', 1):
        ms = timeit(lambda: m.estimate_from_y(r, snr, N, n_summands_or_proba=mode), reps=int(os.environ.get('REPS', 5)))
        print(json.dumps(dict(config=f'GMM full 1-bit N={N} K={K} mode={mode}', B=B, ms=ms, est_per_s=B / ms * 1e3,
                              tflops_algorithmic=16 * K * N * N * B / ms / 1e9)), flush=True)


if __name__ == '__main__':
    main()
