#!/usr/bin/env python
"""Per-pilot error statistics of circ_tc_kernel against the complex128 circ_kernel at config 3 (same pilots, same parameters)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import quantized_channel_estimation_b200 as qce
from quantized_channel_estimation_b200 import synthetic
K, N, B, snr = 128, 256, 4096, 10
c, covs, w, F = synthetic.circulant_gmm(K, 16, 16, seed=0, dense=False)
qz = qce.get_quantizer([snr], 3, 'lloyd')[snr]
m = qce.Gmm_nbit(n_components=K, covariance_type='block-circulant')
m.set_circulant_parameters(c, w, (16, 16))
g = torch.Generator(device='cuda').manual_seed(3)
y = torch.view_as_complex(torch.randn((B, N, 2), generator=g, device='cuda', dtype=torch.float64)) * 0.8
r = qce.quant(y, 3, qz[0], qz[1])
model = m._prepared(np.eye(N), snr, 3, 'lloyd', qz)
e_tc, lp_tc = model.estimate(r, 'all', 'tc', want_logp=True)
e_64, lp_64 = model.estimate(r, 'all', 'fp64', want_logp=True)
per = (e_tc - e_64).norm(dim=1) / e_64.norm(dim=1)
print('est relerr total', float((e_tc - e_64).norm() / e_64.norm()), 'per-pilot max', float(per.max()), 'median', float(per.median()))
d = lp_tc - lp_64
print('logp abs err max', float(d.abs().max()), 'rms', float(d.pow(2).mean().sqrt()))
# error of differences to the per-pilot max component
mx = lp_64.argmax(1)
dd = d - d.gather(1, mx[:, None])
top = (lp_64 - lp_64.max(1, keepdim=True).values) > -15
print('relative-to-max logp err (competitive comps): max', float(dd[top].abs().max()), 'rms', float(dd[top].pow(2).mean().sqrt()))
w64 = torch.softmax(lp_64, 1); wtc = torch.softmax(lp_tc, 1)
print('weight abs err max', float((w64 - wtc).abs().max()))
print('logp magnitude', float(lp_64.abs().mean()), 'spread of top', float((lp_64.max(1).values - lp_64.median(1).values).mean()))
