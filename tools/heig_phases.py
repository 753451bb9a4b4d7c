"""Eigensolver of the truncation at the benchmark sizes: wall time of cb200_k_heig (H2D/D2H excluded by subtracting a copy-only call is
not possible through this ABI, so the matrix upload is reported separately) for the one-stage and the two-stage reduction, plus a
residual check that needs no CPU eigensolver.   python tools/heig_phases.py n m [c]   (CB200_TWOSTAGE_MIN_N selects the path)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chain_b200 as cb

n = int(sys.argv[1]); m = int(sys.argv[2]); cplx = len(sys.argv) > 3 and sys.argv[3] == "c"
rng = np.random.default_rng(0)
X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
X *= np.exp(-np.arange(n) * (14.0 / n))[None, :]
A = X.conj().T @ X
A /= np.trace(A).real
ctx = cb.Context(0)
for mode, thr in (("two-stage", "512"), ("one-stage", "0")):
    if n > 6000 and mode == "one-stage" and os.environ.get("HEIG_SKIP_ONE"):
        continue
    os.environ["CB200_TWOSTAGE_MIN_N"] = thr
    cb.k_heig(A[:2000, :2000].copy(), 100, ctx=ctx)          # warm the pool / code
    t0 = time.perf_counter()
    w, V, it = cb.k_heig(A, m, ctx=ctx)
    t1 = time.perf_counter()
    R = A @ V - V @ (V.conj().T @ A @ V)
    print("%s n=%d m=%d cplx=%d wall %.1f ms ns_iters=%d orth=%.1e resid=%.1e trace_err=%.1e" % (
        mode, n, m, cplx, (t1 - t0) * 1e3, it, np.abs(V.conj().T @ V - np.eye(m)).max(), np.abs(R).max(), abs(w.sum() - 1.0)), flush=True)
    if mode == "two-stage":
        w2 = w
    else:
        print("   max |w_two - w_one| = %.2e" % np.abs(w2 - w).max())
