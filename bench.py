#!/usr/bin/env python
"""bench.py -- Bussgang-GMM estimates/sec (N=64 antennas, K=64 components, 1-bit) on B200.

One "step" = one pass of the hot path (observe -> 1-bit quantise -> Bussgang-GMM 'all' estimate -> NMSE
accumulators) over one batch of 2^20 synthetic observations per GPU at one SNR of the sweep -10..30 dB
(BASELINE.json configs[1]); the per-SNR component parameters are precomputed and resident.

  value      whole-job estimates/s with channels + noise already in HBM (CUDA events, max over ranks)
  e2e        the same metric through Gmm_nbit.estimate_from_y on HOST (pinned) complex128 pilots:
             host->device and device->host copies inside the timed region
  roofline   the dense estimate kernel timed alone: algorithmic 16*K*N^2 flop per estimate
  cpu_baseline  the numpy oracle (a port of the reference algorithm) on the host cores, bounded sample

`--impl reference` times that CPU port alone (the Python reference cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ANT, N_COMP, N_BITS = 64, 64, 1
SNRS = list(range(-10, 31, 5))
BATCH = 1 << 20
METRIC = 'Bussgang-GMM estimates/sec (M=64,K=64,1-bit)'
UNIT = 'estimates/s'
FLOP_PER_EST = 16 * N_COMP * N_ANT * N_ANT
WORKLOAD = f"Bussgang-GMM 'full' 1-bit N={N_ANT} K={N_COMP} mode=all, SNR sweep -10..30 dB (BASELINE configs[1])"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(burst=float(p['bf16_tflops']), sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                    hbm=float(p['hbm_gbs']), src='measured')
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_params():
    from quantized_channel_estimation_b200 import synthetic      # seeded synthetic generators (SURVEY.md section 8d P-rand)
    return synthetic.random_psd_gmm(N_COMP, N_ANT, seed=0)


def cpu_port_rate(means, covs, w, n_obs, snr=10, seed=123):
    """Time the numpy oracle (port of gmm:166-243) on n_obs observations; returns (est/s, seconds)."""
    from oracle import qce_oracle as orc
    h, noise, _ = orc.sample_gmm_channels(means, covs, w, n_obs, seed=seed)
    t0 = time.perf_counter()
    r = orc.get_observation_nbit(h, snr, noise, None, N_BITS)
    est = orc.gmm_estimate_from_y(means, covs, w, r, snr, n_summands_or_proba='all', n_bits=N_BITS)
    dt = time.perf_counter() - t0
    return n_obs / dt, dt, float(orc.mse(est, h))


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get('num_threads', 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    use_all_host_threads()
    means, covs, w = make_params()
    sample = 4096
    for _ in range(args.warmup):
        cpu_port_rate(means, covs, w, 256)
    t = 0.0
    for i in range(args.steps):
        _, dt, _ = cpu_port_rate(means, covs, w, sample, snr=SNRS[i % len(SNRS)], seed=1000 + i)
        t += dt
    val = sample * args.steps / t
    cores = blas_threads()
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'c128', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_step': sample, 'note': 'bounded sample of the same workload on the host cores'},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': f'{sample} observations per step (numpy oracle, batched over samples), {os.cpu_count()} host cpus'},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=9)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--precision', default='auto', choices=['auto', 'tc', 'fp64'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import quantized_channel_estimation_b200 as qce
    from quantized_channel_estimation_b200 import _lib, engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _lib.require_device()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    means, covs, w = make_params()
    gmm = qce.Gmm_nbit(n_components=N_COMP, covariance_type='full').set_parameters(means, covs, w, zero_mean=True)
    gmm.precision = args.precision
    eye = np.eye(N_ANT, dtype=complex)
    models = [gmm._prepared(eye, s, N_BITS, 'uniform', None) for s in SNRS]          # resident per-SNR parameters
    quant = engine.Quantizer.get(N_BITS)
    B = args.batch

    # synthetic data, generated on the device (seeded by rank): h = C_k^{1/2} g, k ~ Cat(w), complex64 like SCMMulti
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    Lc = torch.linalg.cholesky(torch.as_tensor(covs, device=dev))
    lab = torch.multinomial(torch.as_tensor(w, device=dev), B, replacement=True, generator=g)
    h = torch.empty((B, N_ANT), dtype=torch.complex64, device=dev)
    for k in range(N_COMP):
        idx = (lab == k).nonzero(as_tuple=True)[0]
        gk = torch.randn((idx.numel(), N_ANT, 2), generator=g, device=dev, dtype=torch.float64)
        h[idx] = (torch.view_as_complex(gk) * np.sqrt(0.5) @ Lc[k].T).to(torch.complex64)
    noise = torch.view_as_complex(torch.randn((B, N_ANT, 2), generator=g, device=dev, dtype=torch.float64)) * np.sqrt(0.5)
    noise = noise.contiguous()
    acc = torch.zeros((len(SNRS), 3), dtype=torch.float64, device=dev)
    acc_sweep, acc_total = torch.zeros_like(acc), torch.zeros_like(acc)
    del Lc, lab

    def step(i):
        j = i % len(SNRS)
        models[j].pipeline(quant, h, noise, 10 ** (-SNRS[j] / 20), 'all', args.precision, acc=acc[j])
        # the path's only exchange: ONE all-reduce of the [n_snr, 3] NMSE accumulators per completed SNR sweep
        if world > 1 and j == len(SNRS) - 1:
            dist.all_reduce(acc_sweep.copy_(acc))
            acc_total.add_(acc_sweep)
            acc.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, len(SNRS))):      # every resident per-SNR model is exercised before the clock starts
        step(i)
    barrier()
    acc.zero_()
    acc_total.zero_()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    if world > 1:                                        # remainder of an unfinished sweep (inside the timed region)
        dist.all_reduce(acc_sweep.copy_(acc))
        acc_total.add_(acc_sweep)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)
    accs = (acc_total if world > 1 else acc).cpu().numpy()

    # --- roofline: the dense estimate kernel alone on resident quantised pilots
    pk = peaks()
    import ctypes as C
    lib = _lib.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    r = qce.get_observation_nbit(h, SNRS[4], n_bits=N_BITS, noise=noise)
    out = torch.empty((B, N_ANT), dtype=torch.complex128, device=dev)
    tc_used = args.precision in ('auto', 'tc') and lib.qce_format_pilots(models[4].handle, stream, C.c_void_p(r.data_ptr()), B) == 0
    if tc_used:
        def kernel_only():
            _lib.check(lib.qce_estimate_formatted(models[4].handle, stream, B, C.c_void_p(out.data_ptr()), None, None))
    else:
        def kernel_only():
            models[4].estimate(r, 'all', 'fp64')
    for _ in range(2):
        kernel_only()
    torch.cuda.synchronize()
    reps = 5
    ev0.record()
    for _ in range(reps):
        kernel_only()
    ev1.record()
    torch.cuda.synchronize()
    k_ms = ev0.elapsed_time(ev1) / reps
    achieved = FLOP_PER_EST * B / (k_ms * 1e-3) / 1e12
    # DRAM traffic of one 2^20-pilot launch of the dominant kernel from the committed ncu --set full capture
    # (profiles/r01_tc_final_ncu_summary.txt: dram read 295.9 MB + write 1032.4 MB; algorithmic: 268 MB tiles + 1074 MB estimates)
    traffic = 1.328e9 * (B / float(1 << 20)) if tc_used else None
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['burst'], 'unit': 'TFLOP/s', 'frac': achieved / pk['burst'],
                'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)', 'kernel': 'dense_tc_kernel' if tc_used else 'dense_fp64_kernel', 'kernel_ms': k_ms,
                'peak_source': f"{pk['src']} bf16 burst (kernel timed alone)",
                'algorithmic_flop_per_estimate': FLOP_PER_EST,
                'tensor_passes': 2 if tc_used else None,
                'note': ('FP16 hi/lo split: 2 tensor passes per algorithmic flop; cta_group::2 MMAs' if tc_used else
                         'complex128 SIMT kernel: bounded by the FP64 pipe (~40 TFLOP/s), not the tensor pipe')}
    del r, out

    # --- e2e: host pinned complex128 pilots -> estimate_from_y -> host estimates
    e2e = None
    if not args.no_e2e:
        Be = B
        r_host = torch.empty((Be, N_ANT), dtype=torch.complex128).pin_memory()
        r_host.copy_(qce.get_observation_nbit(h[:Be], SNRS[4], n_bits=N_BITS, noise=noise[:Be]).cpu())
        r_np = r_host.numpy()
        gmm.estimate_from_y(r_np[:4096], SNRS[4], N_ANT, n_summands_or_proba='all', n_bits=N_BITS)
        from quantized_channel_estimation_b200.engine import parse_mode
        out_host = torch.empty((Be, N_ANT), dtype=torch.complex128).pin_memory()
        mode, n_top, rho = parse_mode('all')
        prec = _lib.PREC_TC if tc_used else _lib.PREC_FP64

        def e2e_step(j):
            _lib.check(lib.qce_estimate_host(models[j].handle, C.c_void_p(r_host.data_ptr()), Be, mode, n_top, rho, prec,
                                             C.c_void_p(out_host.data_ptr())))
        for i in range(2):
            e2e_step(i % len(SNRS))
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i % len(SNRS))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {'value': world * Be * args.steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': Be * N_ANT * 16,
               'd2h_bytes_per_step': Be * N_ANT * 16, 'api': 'qce_estimate_host (Gmm_nbit.estimate_from_y on host arrays)',
               'host_buffers': 'pinned'}
        del r_host, out_host
    clk = clocks.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        n_cpu = 8192
        cpu_port_rate(means, covs, w, 256)
        rate, dt, _ = cpu_port_rate(means, covs, w, n_cpu)
        cpu = {'value': rate, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
               'sample': f'{n_cpu} observations at 10 dB, numpy oracle batched over samples ({dt:.1f} s), {os.cpu_count()} host cpus'}

    if rank == 0:
        nmse = {str(s): float(a[0] / max(a[2], 1) / N_ANT) for s, a in zip(SNRS, accs) if a[2] > 0}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f16x2-split/f32-acc' if tc_used else 'f64', 'data': 'synthetic',
                'config': {'workload': WORKLOAD,
                           'batch_per_gpu_per_step': B, 'global_batch': world * B, 'parallelism': f'dp{world}',
                           'l2': 'inputs larger than L2 (h 512 MiB c64 + noise 1 GiB c128 per step), no flush needed',
                           'params': 'random-PSD GMM seed 0 (SURVEY 8d P-rand)'},
                'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clk,
                'nmse_per_snr': nmse}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
