"""World-size-2 CPU test (gloo) of the sample-sharded sweep: contiguous shards, one all-reduce of the NMSE
accumulators, result equal to the single-process sweep.  The per-SNR step is the numpy oracle here (the checker
standing in for the GPU kernels, which cannot run on the CPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from quantized_channel_estimation_b200.montecarlo import nmse_sweep, shard_range

SNRS = [-5, 10]
K, N, B = 4, 8, 101


def _data():
    from oracle import qce_oracle as orc
    means, covs, w = orc.random_psd_gmm(K, N, seed=2)
    h, noise, _ = orc.sample_gmm_channels(means, covs, w, B, seed=3)
    return orc, means, covs, w, h, noise


def _step_factory():
    orc, means, covs, w, h, noise = _data()

    def step(i, snr, lo, hi):
        r = orc.get_observation_nbit(h[lo:hi], snr, noise[lo:hi], None, 1)
        est = orc.gmm_estimate_from_y(means, covs, w, r, snr, n_summands_or_proba='all', n_bits=1)
        return (np.sum(np.abs(est - h[lo:hi]) ** 2), np.sum(np.abs(h[lo:hi]) ** 2), hi - lo)
    return step


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    nmse, acc = nmse_sweep(_step_factory(), B, SNRS, N)
    q.put((rank, nmse, acc.numpy()))
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (0, 1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sweep_matches_single_process():
    single, acc1 = nmse_sweep(_step_factory(), B, SNRS, N)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nmse, acc in res:
        np.testing.assert_allclose(nmse, single, rtol=1e-12)
        np.testing.assert_allclose(acc, acc1.numpy(), rtol=1e-12)
        assert acc[0, 2] == B
