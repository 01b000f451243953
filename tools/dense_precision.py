#!/usr/bin/env python
"""Per-pilot error statistics of the tensor-core paths against the complex128 kernel (same pilots, same parameters):
total relative error, worst pilot, error of l_k - l_max over the competitive components."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantized_channel_estimation_b200 as qce
from quantized_channel_estimation_b200 import engine, precompute, synthetic


def stats(tag, model, r):
    e_tc, lp_tc = model.estimate(r, 'all', 'tc', want_logp=True)
    e_64, lp_64 = model.estimate(r, 'all', 'fp64', want_logp=True)
    per = (e_tc - e_64).norm(dim=1) / e_64.norm(dim=1)
    d = lp_tc - lp_64
    dd = d - d.gather(1, lp_64.argmax(1)[:, None])
    top = (lp_64 - lp_64.max(1, keepdim=True).values) > -15
    print(json.dumps(dict(case=tag, est_relerr_total=float((e_tc - e_64).norm() / e_64.norm()), est_relerr_worst_pilot=float(per.max()),
                          est_relerr_median_pilot=float(per.median()), logp_rel_to_max_rms=float(dd[top].pow(2).mean().sqrt()),
                          logp_rel_to_max_max=float(dd[top].abs().max()), logp_magnitude=float(lp_64.abs().mean()))), flush=True)


def main():
    B = 8192
    for tag, K, N, snr, nb, qt in (('C2 N=64 K=64 1-bit 10dB', 64, 64, 10, 1, 'uniform'), ('C2 -10dB', 64, 64, -10, 1, 'uniform'),
                                   ('C2 30dB', 64, 64, 30, 1, 'uniform'), ('N=128 K=64 2-bit split path', 64, 128, 10, 2, 'uniform'),
                                   ('N=64 K=32 3-bit Lloyd (off-grid)', 32, 64, 10, 3, 'lloyd')):
        means, covs, w = synthetic.random_psd_gmm(K, N, seed=0)
        h, noise, _ = synthetic.sample_gmm_channels(means, covs, w, B, seed=1)
        qz = qce.get_quantizer([snr], nb, qt)[snr]
        r = qce.get_observation_nbit(torch.from_numpy(h).cuda(), snr, n_bits=nb, thresholds=qz[0], cluster=qz[1], noise=torch.from_numpy(noise).cuda())
        model = engine.DenseModel(precompute.prepare(means, covs, w, np.eye(N), snr, nb, qt, qz))
        stats(tag, model, r)


if __name__ == '__main__':
    main()
