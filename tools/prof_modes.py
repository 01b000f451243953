#!/usr/bin/env python
"""One call of a hard-selection mode at the config-2 / config-4 shape on model-drawn pilots: the launch list under
`ncu --metrics gpu__time_duration.sum` shows where the time of the whitening -> selection -> (pair-)bucketed combine path goes.
    SHAPE=c2|c4 MODE=4|0.9|1|all python tools/prof_modes.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import quantized_channel_estimation_b200 as qce
from quantized_channel_estimation_b200 import synthetic
from bench_configs import model_pilots, timeit


def main():
    shape, mode = os.environ.get('SHAPE', 'c2'), os.environ.get('MODE', '4')
    mode = 'all' if mode == 'all' else (float(mode) if '.' in mode else int(mode))
    B, snr = 1 << int(os.environ.get('LOG2B', 19)), int(os.environ.get('SNR', 10))
    if shape == 'c2':
        means, covs, w = synthetic.random_psd_gmm(64, 64, seed=0)
        m = qce.Gmm_nbit(n_components=64).set_parameters(means, covs, w, detect_structure=False)
        r = model_pilots(covs, w, B, snr, 1, (None, None))
        fn = lambda: m.estimate_from_y(r, snr, 64, n_summands_or_proba=mode)
    else:
        means, lambdas, psis, amps = synthetic.random_mfa(64, 128, 16, seed=0)
        qz = qce.get_quantizer([snr], 2, 'uniform')[snr]
        m = qce.Mofa(64, 16, verbose=False).set_parameters(means, lambdas, psis, amps)
        m.use_structure = False
        m.precision = 'tc'
        r = model_pilots(m.covs, amps, B, snr, 2, qz)
        fn = lambda: m.estimate_from_y(r, snr, n_summands_or_proba=mode, n_bits=2, quantizer_type='uniform', quantizer=qz)
    ms = timeit(fn, reps=int(os.environ.get('REPS', 3)))
    print(json.dumps(dict(shape=shape, mode=mode, B=B, ms=ms, est_per_s=B / ms * 1e3)))


if __name__ == '__main__':
    main()
