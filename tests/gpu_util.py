"""Helpers for the -m gpu tests: device buffers are plain torch tensors (torch is plumbing only)."""
import numpy as np
import torch


def dev(buf):
    return torch.from_numpy(np.ascontiguousarray(buf)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def gemm_tolerance(dtype, k, absab, exact):
    """BASELINE.json: |C - C_ref| <= c*k*eps*(|A||B|), c = 2 (FP64, eps = 2^-52) or c = 8 (3xTF32 /
    FP32, eps = 2^-23), plus one rounding of the result itself."""
    eps = np.finfo(dtype).eps
    c = 2 if np.dtype(dtype) == np.float64 else 8
    return c * max(k, 1) * eps * absab + eps * np.abs(exact)
