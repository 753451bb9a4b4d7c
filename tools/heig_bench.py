"""Standalone run of the truncation eigensolver (cb200_k_heig) at a given size; meant to be run under
`ncu --metrics gpu__time_duration.sum` so that the per-kernel device times can be read from the launch list.

    python tools/heig_bench.py n m [c] [batch]     # batch > 1: `batch` matrices of that size through cb200_k_heig_batched (the charge
                                                   # blocks of one bond side by side; CB200_TRI_BATCH=0 for the sequential A/B)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chain_b200 as cb

n = int(sys.argv[1]); m = int(sys.argv[2]); cplx = len(sys.argv) > 3 and sys.argv[3] == "c"
rng = np.random.default_rng(0)
X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
X *= np.exp(-np.arange(n) * (12.0 / n))[None, :]
A = X.conj().T @ X
A /= np.trace(A).real
ctx = cb.Context(0)
cb.k_heig(A[:64, :64].copy(), 8, ctx=ctx)
t0 = time.perf_counter()
w, V, it = cb.k_heig(A, m, ctx=ctx)
t1 = time.perf_counter()
print("n=%d m=%d cplx=%d wall %.1f ms (incl. H2D/D2H) ns_iters=%d orth=%.1e" % (n, m, cplx, (t1 - t0) * 1e3, it, np.abs(V.conj().T @ V - np.eye(m)).max()))
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
if batch > 1:
    As = [A] + [A + 1e-3 * k * np.eye(n) for k in range(1, batch)]
    cb.k_heig_batched([a[:64, :64].copy() for a in As], [8] * batch, ctx=ctx)
    t0 = time.perf_counter()
    out = cb.k_heig_batched(As, [m] * batch, ctx=ctx)
    t1 = time.perf_counter()
    print("batch of %d: wall %.1f ms (incl. H2D/D2H) max|w - w_single| = %.1e" % (batch, (t1 - t0) * 1e3, np.abs(out[0][0] - w).max()))
