"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel of the library once."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qce_oracle as orc
import quantized_channel_estimation_b200 as qce

K, N, B, snr = 8, 64, 700, 10
means, covs, w = orc.random_psd_gmm(K, N, seed=0)
h, noise, _ = orc.sample_gmm_channels(means, covs, w, B, seed=1)
m = qce.Gmm_nbit(n_components=K).set_parameters(means, covs, w, detect_structure=False)
r = qce.get_observation_nbit(torch.from_numpy(h).cuda(), snr, n_bits=1, noise=torch.from_numpy(noise).cuda())
ref = orc.gmm_estimate_from_y(means, covs, w, r.cpu().numpy(), snr, n_summands_or_proba='all', n_bits=1)
for prec in ('tc', 'fp64'):
    m.precision = prec
    for mode in ('all', 1, 3, 0.9):
        est = m.estimate_from_y(r, snr, N, n_summands_or_proba=mode)
    est = m.estimate_from_y(r, snr, N, n_summands_or_proba='all').cpu().numpy()
    print(prec, np.linalg.norm(est - ref) / np.linalg.norm(ref))
c, ccovs, cw, _ = orc.circulant_gmm(4, 4, 8, seed=1)
mc = qce.Gmm_nbit(n_components=4).set_circulant_parameters(c, cw, (4, 8))
mc.estimate_from_y(qce.quant(torch.from_numpy(orc.crandn(100, 32, rng=np.random.default_rng(0))).cuda(), 1), snr, 32, n_summands_or_proba='all')
mm, lam, psi, amps = orc.random_mfa(3, 24, 4, seed=2)
mf = qce.Mofa(3, 4, verbose=False).set_parameters(mm, lam, psi, amps)
qz = qce.get_quantizer([snr], 2, 'uniform')[snr]
mf.estimate_from_y(qce.quant(torch.from_numpy(orc.crandn(100, 24, rng=np.random.default_rng(0))).cuda(), 2, qz[0], qz[1]), snr, n_summands_or_proba='all',
                   n_bits=2, quantizer=qz)
torch.cuda.synchronize()
print('done')
