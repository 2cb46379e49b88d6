import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
for (m, n, lda, ldc) in [(517, 300, 301, 544), (517, 300, 301, 544), (256, 256, 257, 256)]:
    rng = np.random.default_rng(1)
    A = rng.uniform(-1, 1, lda * m)
    C0 = np.full(ldc * n, np.nan)
    dA, dC = torch.from_numpy(A).cuda(), torch.from_numpy(C0).cuda()
    h.geam(True, 1, 0, m, n, 1.0, dA, lda, 0.0, dC, ldc, dC, ldc)
    torch.cuda.synchronize()
    got = dC.cpu().numpy().reshape(n, ldc)          # got[j, i] = C(i, j)
    want = np.full((n, ldc), np.nan)
    want[:, :m] = A.reshape(m, lda)[:, :n].T
    bad = np.argwhere(got.view(np.uint64) != want.view(np.uint64))
    print(m, n, lda, ldc, h.last_kernel(), "mismatches", bad.shape[0])
    for jj, ii in bad[:16]:
        print("   C(i=%d, j=%d) got %r want %r  (A flat index of got value: %s)" % (ii, jj, got[jj, ii], want[jj, ii], np.argwhere(A == got[jj, ii]).ravel()[:3]))
    if bad.shape[0]:
        print("   i range", bad[:, 1].min(), bad[:, 1].max(), "j range", bad[:, 0].min(), bad[:, 0].max())
