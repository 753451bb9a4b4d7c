"""Oracle (test infrastructure): CPU timing of the reference's contraction sequence for the H_eff matvec.

The reference's own CPU path is ITensor (QDense contraction = for every compatible block pair: explicit permute
copy + zgemm/dgemm) on OpenBLAS/MKL; it cannot be built here, so this is a *port*: the same
block-pair GEMM list with the same shapes, run through NumPy/OpenBLAS with all host threads.  It times a bounded
SAMPLE of the block pairs of the two D^3 stages (phi*L and (.)*R) of one matvec and reports the achieved rate;
bench.py states the sample next to the number.
"""
import os
import time

import numpy as np


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def block_pair_list(Dl, Dr, kl, kr, ds):
    """GEMM shapes (M, N, K) of ITensor's block-pair loop for stage 1 (L x phi) and stage 4 (. x R), charge-fused site blocks."""
    nq = len(Dl)
    shapes = []
    for ql in range(nq):
        ncols = sum(ds[q1] * ds[q2] * Dr[(ql + q1 + q2) % nq] for q1 in range(nq) for q2 in range(nq))
        for qa in range(nq):
            shapes.append(("L-stage", Dl[(ql + qa) % nq] * kl[qa], ncols, Dl[ql]))
    for qlp in range(nq):
        for qrp in range(nq):
            nrow = Dl[qlp] * sum(ds[q1] * ds[q2] for q1 in range(nq) for q2 in range(nq) if (qlp + q1 + q2) % nq == qrp)
            for qa in range(nq):
                shapes.append(("R-stage", nrow, Dr[qrp], kr[qa] * Dr[(qrp - qa) % nq]))
    return shapes


def matvec_sample(Dl, Dr, kl, kr, ds, cplx, budget_s=15.0, seed=0):
    """Run block-pair GEMMs (with ITensor's explicit permute copy of one operand) until the time budget is used.
    Returns dict(tflops, seconds, flops, sample)."""
    rng = np.random.default_rng(seed)
    shapes = block_pair_list(Dl, Dr, kl, kr, ds)
    order = np.arange(len(shapes))
    rng.shuffle(order)
    flops, secs, done = 0.0, 0.0, 0
    per = 8.0 if cplx else 2.0
    for i in order:
        _, M, N, K = shapes[i]
        if M == 0 or N == 0 or K == 0:
            continue
        A = rng.standard_normal((K, M)) + (1j * rng.standard_normal((K, M)) if cplx else 0)   # stored transposed: needs the permute copy
        B = rng.standard_normal((K, N)) + (1j * rng.standard_normal((K, N)) if cplx else 0)
        t0 = time.perf_counter()
        At = np.ascontiguousarray(A.T)        # ITensor permutes the operand into matrix form before calling gemm
        C = At @ B
        secs += time.perf_counter() - t0
        flops += per * M * N * K
        done += 1
        del A, B, At, C
        if secs >= budget_s:
            break
    return dict(tflops=flops / secs / 1e12, seconds=secs, flops=flops,
                sample="%d of %d block-pair GEMMs (permute copy + %s) of one matvec's two D^3 stages" % (done, len(shapes), "zgemm" if cplx else "dgemm"))
