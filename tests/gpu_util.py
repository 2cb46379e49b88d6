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


def sampled_exact_ratio(oracle, dtype, ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc, samples=24, seed=7, alpha=1.0):
    """max over a sampled grid of entries (corners included) of |C - C_exact| / (c k eps (|A||B|) + eps |C_exact|), the bound of
    BASELINE.json at 1x, with C_exact and |A||B| from the oracle's long-double product of the gathered rows of op(A) and
    columns of op(B).  A, B, Cm are device tensors holding column-major matrices with the given leading dimensions (beta = 0)."""
    g = torch.Generator().manual_seed(seed)
    rows = torch.unique(torch.cat([torch.tensor([0, m - 1]), torch.randint(0, m, (samples,), generator=g)])).cuda()
    cols = torch.unique(torch.cat([torch.tensor([0, n - 1]), torch.randint(0, n, (samples,), generator=g)])).cuda()
    ac = m if ta else k
    bc = k if tb else n
    Av = A[: lda * ac].view(ac, lda)              # stored element (r, c) = Av[c, r]
    sub_a = Av[rows, :k] if ta else Av[:k, rows].t()          # (r, k): rows of op(A)
    Bv = B[: ldb * bc].view(bc, ldb)
    sub_b = Bv[:k, cols] if tb else Bv[cols, :k].t()          # (k, c): columns of op(B)
    r, c = rows.numel(), cols.numel()
    ha = np.ascontiguousarray(sub_a.t().cpu().numpy())        # (k, r) C-contiguous == (r x k) column-major, ld r
    hb = np.ascontiguousarray(sub_b.t().cpu().numpy())        # (c, k) C-contiguous == (k x c) column-major, ld k
    exact, absab = oracle.gemm_exact(0, 0, r, c, k, alpha, ha.reshape(-1), r, hb.reshape(-1), k, 0.0, None, r)
    got = Cm[: ldc * n].view(n, ldc)[cols][:, rows].t().cpu().numpy().astype(np.float64)
    return float(np.max(np.abs(got - exact) / gemm_tolerance(dtype, k, absab, exact)))
