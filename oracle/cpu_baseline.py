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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """Set the BLAS pool to every host core this process may run on and return the thread count ACTUALLY in use.
    torchrun exports OMP_NUM_THREADS=1 to its workers, which silently pinned OpenBLAS to one thread in round 1 while the line still
    said `cores: 32`; the pool size is therefore set explicitly (threadpoolctl -> openblas_set_num_threads) and read back."""
    from threadpoolctl import threadpool_info, threadpool_limits
    threadpool_limits(limits=host_cores(), user_api="blas")
    n = [i["num_threads"] for i in threadpool_info() if i.get("user_api") == "blas"]
    return max(n) if n else 1


def host_threads():
    return use_all_host_threads()


def eigh_seconds(n, cplx, seed=0):
    """Wall time of LAPACK *syev / *heev (the driver ITensor's diagHermitian calls) on an n x n density-matrix-like input."""
    import scipy.linalg as sla
    use_all_host_threads()
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
    A = X @ X.conj().T
    A /= np.trace(A).real
    t0 = time.perf_counter()
    sla.eigh(A, driver="ev", check_finite=False)
    return time.perf_counter() - t0


def sweep_estimate(matvec_flops_total, env_flops_total, block_sizes, cplx, gemm_tflops, n_meas_cap=None):
    """CPU seconds of one sweep of the SAME workload the GPU sweep ran: the sweep's contraction flops at the CPU GEMM rate measured by
    matvec_sample, plus one *syev / *heev per density-matrix block, timed at min(largest block, cap) and scaled with n^3.
    block_sizes: the orders of all rho blocks of all bond updates of the sweep.  Returns dict(seconds, parts, how)."""
    nmax = max(block_sizes) if block_sizes else 0
    cap = n_meas_cap or (1200 if cplx else 1600)
    nm = max(2, min(nmax, cap))
    t_m = eigh_seconds(nm, cplx)
    t_eig = sum(t_m * (n / nm) ** 3 for n in block_sizes)
    t_gemm = (matvec_flops_total + env_flops_total) / (gemm_tflops * 1e12)
    return dict(seconds=t_gemm + t_eig, contraction_s=t_gemm, eigensolver_s=t_eig,
                how="EXTRAPOLATED: (matvec + environment flops of the sweep) / the CPU block-pair GEMM rate measured above + one LAPACK %s per "
                    "rho block, timed at n=%d (%.2f s) and scaled with n^3 over the %d blocks of the sweep (largest n=%d)"
                    % ("zheev" if cplx else "dsyev", nm, t_m, len(block_sizes), nmax))


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
    use_all_host_threads()
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
