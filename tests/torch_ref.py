"""Torch float64 composition of the per-sample maths on the precomputed parameter blocks (GEMM +
softmax).  TEST HELPER: validates the host precompute/packing against the reference golden vectors on
the CPU, and serves as the fp64 torch reference for the floating-point CUDA kernels."""
import numpy as np
import torch


def reference_combine(prep, r, mode='all', top1_exp_argmax=False):
    r = torch.as_tensor(np.asarray(r), dtype=torch.complex128, device=prep['Linv'].device)
    z = torch.einsum('kij,bj->bki', prep['Linv'], r) - prep['zoff'][None]
    lp = prep['logc'][None] - (z.real ** 2 + z.imag ** 2).sum(-1)
    hk = torch.einsum('knj,bj->bkn', prep['W'], r) + prep['hoff'][None]
    p = torch.softmax(lp, dim=1)
    if mode == 'all':
        wts = p
    elif isinstance(mode, int) and mode == 1:
        lab = lp.argmax(1)
        if top1_exp_argmax:
            lab = torch.exp(lp).argmax(1)
        wts = torch.nn.functional.one_hot(lab, lp.shape[1]).to(p.dtype)
    else:
        ps, idx = torch.sort(p, dim=1, descending=True)
        if isinstance(mode, int):
            keep = torch.arange(p.shape[1], device=p.device)[None, :] < mode
        else:
            cs = torch.cumsum(ps, 1)
            nr = (cs < mode).sum(1) + 1
            keep = torch.arange(p.shape[1], device=p.device)[None, :] < nr[:, None]
        sel = ps * keep
        wts = torch.zeros_like(p).scatter(1, idx, sel / sel.sum(1, keepdim=True))
    return torch.einsum('bk,bkn->bn', wts.to(hk.dtype), hk), lp
