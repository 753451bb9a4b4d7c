"""-m gpu: the device kernels through the C ABI, against NumPy on the same seeded inputs (bit-tolerance 1e-12..1e-13
relative: FP64 accumulation order differs from BLAS)."""
import numpy as np
import pytest

import chain_b200 as cb
import common

pytestmark = pytest.mark.gpu


def _rand(rng, shape, cplx):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if cplx else x


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (257, 131, 95), (1, 1, 1), (70, 45, 33), (533, 300, 534), (64, 640, 7)])
def test_gemm_single_segment(gpu_ctx, cplx, ta, tb, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K + cplx)
    A = _rand(rng, (M, K), cplx)
    B = _rand(rng, (K, N), cplx)
    a = (A.T if ta else A).ravel(order="F")      # ta: stored K x M column-major
    b = (B.T if tb else B).ravel(order="F")
    lda = K if ta else M
    ldb = N if tb else K
    c, _ = cb.k_gemm(a, b, np.zeros(M * N), M, N, [K], [0], [0], lda, ldb, 0, M, trans_a=ta, trans_b=tb, ctx=gpu_ctx)
    want = A @ B
    assert np.abs(c.reshape(N, M).T - want).max() <= 1e-13 * K * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("cplx", [False, True])
def test_gemm_conj_accumulate_offsets_segments(gpu_ctx, cplx):
    rng = np.random.default_rng(5 + cplx)
    M, N = 150, 77
    Ks = [33, 0, 129, 5]
    lda, ldb, ldc = 161, 140, 157           # odd leading dimensions + odd offsets: exercises the unaligned (8-byte) load path
    A = _rand(rng, lda * (M + 8), cplx)
    B = _rand(rng, ldb * N + 50, cplx)
    C0 = _rand(rng, ldc * N + 9, cplx)
    a_off, b_off = [3, 0, 11, 1000], [1, 0, 7, 2]
    c, _ = cb.k_gemm(A, B, C0, M, N, Ks, a_off, b_off, lda, ldb, 5, ldc, trans_a=True, conj_a=True, accumulate=True, ctx=gpu_ctx)
    want = C0.copy()
    Cm = np.zeros((M, N), dtype=want.dtype)
    for ao, bo, k in zip(a_off, b_off, Ks):
        if k == 0:
            continue
        Aseg = np.array([[A[ao + kk + m * lda] for kk in range(k)] for m in range(M)])     # element (m,k) at k + m*lda
        Bseg = np.array([[B[bo + kk + n * ldb] for n in range(N)] for kk in range(k)])
        Cm += Aseg.conj() @ Bseg
    for n in range(N):
        want[5 + n * ldc: 5 + n * ldc + M] += Cm[:, n]
    assert np.abs(c - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
    # untouched padding must be bit-identical
    mask = np.ones(len(want), bool)
    for n in range(N):
        mask[5 + n * ldc: 5 + n * ldc + M] = False
    assert np.array_equal(c[mask], C0[mask])


@pytest.mark.parametrize("cplx", [False, True])
def test_dots_and_lincomb(gpu_ctx, cplx):
    rng = np.random.default_rng(9)
    for n in (1, 1000, 1 << 20 | 3):
        xs = _rand(rng, (5, n), cplx)
        y = _rand(rng, n, cplx)
        d = cb.k_dots(xs, y, ctx=gpu_ctx)
        want = xs.conj() @ y
        assert np.abs(d - want).max() <= 1e-12 * max(1.0, np.abs(want).max()) * np.sqrt(n)
        d2 = cb.k_dots(xs, y, ctx=gpu_ctx)
        assert np.array_equal(d, d2)          # deterministic reduction (replicated Krylov vectors stay bit-identical)
        co = _rand(rng, 5, cplx)
        out = cb.k_lincomb(xs, co, y, beta=0.5, ctx=gpu_ctx)
        assert np.abs(out - (0.5 * y + co @ xs)).max() <= 1e-13 * 10
        out = cb.k_lincomb(xs[:1], co[:1], y, beta=0.0, ctx=gpu_ctx)
        assert np.abs(out - co[0] * xs[0]).max() <= 1e-14 * 10


@pytest.mark.parametrize("cplx", [False, True])
def test_jacobi_eig_kernel(gpu_ctx, cplx):
    rng = np.random.default_rng(3)
    Y = _rand(rng, (7, 200, 64), cplx)
    Y[3, :, 10:] *= 1e-6                 # graded columns
    Y[4, :, 40:] = 0                     # rank deficient
    G = np.einsum("pki,pkj->pij", Y.conj(), Y)
    Q, rot = cb.k_jacobi_eig(G, 1e-14, ctx=gpu_ctx)
    assert rot > 0
    for p in range(7):
        assert np.abs(Q[p].conj().T @ Q[p] - np.eye(64)).max() <= 1e-13
        D = Q[p].conj().T @ G[p] @ Q[p]
        dd = np.sqrt(np.abs(np.outer(np.diag(D).real, np.diag(D).real))) + 1e-300
        off = np.abs(D - np.diag(np.diag(D))) / dd
        ok = (np.abs(np.diag(D))[:, None] > 1e-20 * np.abs(D).max()) & (np.abs(np.diag(D))[None, :] > 1e-20 * np.abs(D).max())
        assert (off[ok]).max() <= 1e-12
        assert np.allclose(np.sort(np.diag(D).real), np.linalg.eigvalsh(G[p]), atol=1e-12 * np.abs(D).max())


@pytest.mark.parametrize("rows,cols", [(90, 70), (30, 70), (64, 64), (5, 3), (1, 1), (600, 300), (300, 500)])
@pytest.mark.parametrize("cplx", [False, True])
def test_block_jacobi_svd(gpu_ctx, rows, cols, cplx):
    rng = np.random.default_rng(rows * 100 + cols)
    X = _rand(rng, (rows, cols), cplx)
    if cols > 5:
        X[:, 5] = 0
    xv, v, w, sweeps = cb.k_svd(X, ctx=gpu_ctx)
    s = np.linalg.svd(X, compute_uv=False)
    k = min(rows, cols)
    assert np.abs(w[:k] - s ** 2).max() <= 1e-11 * max(1.0, s[0] ** 2)
    assert np.abs(v.conj().T @ v - np.eye(cols)).max() <= 1e-11
    assert np.abs(X @ v - xv).max() <= 1e-12 * max(1.0, s[0])
    g = xv.conj().T @ xv
    assert np.abs(g - np.diag(np.diag(g))).max() <= 1e-11 * max(1.0, s[0] ** 2)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,m,decay", [(1, 1, 0.0), (2, 1, 0.5), (3, 3, 0.3), (37, 20, 0.4), (200, 200, 0.0), (333, 100, 0.08), (1000, 300, 0.03)])
def test_hermitian_eigensolver(gpu_ctx, cplx, n, m, decay):
    """rho -> tridiagonal -> bisection -> inverse iteration -> Newton-Schulz -> back-transformation, against numpy.eigh."""
    rng = np.random.default_rng(n * 10 + m + cplx)
    X = _rand(rng, (n + 5, n), cplx) * (10.0 ** (-decay * np.arange(n)))      # graded spectrum like a DMRG rho
    if n > 30:
        X[:, 7] = X[:, 3]                                                      # an exactly rank-deficient direction
        X[:, 11] = 1j * X[:, 4] if cplx else -X[:, 4]
    A = X.conj().T @ X
    A /= np.trace(A).real
    w, V, it = cb.k_heig(A, m, ctx=gpu_ctx)
    ev, evec = np.linalg.eigh(A)
    assert np.abs(w - ev[::-1]).max() <= 1e-14 * np.sqrt(n)
    assert np.abs(V.conj().T @ V - np.eye(m)).max() <= 1e-13
    # V spans an invariant subspace: residual of the Rayleigh-Ritz pair at the eps*||A|| level
    R = A @ V - V @ (V.conj().T @ A @ V)
    assert np.abs(R).max() <= 2e-14 * np.sqrt(n)
    # and it is the TOP-m one wherever that is well defined (gap at the cut well above the noise)
    if m < n and ev[::-1][m - 1] - ev[::-1][m] > 1e-9:
        P = evec[:, n - m:] @ evec[:, n - m:].conj().T
        assert np.abs(V @ V.conj().T - P).max() <= 1e-5
    assert np.abs(np.sort(np.linalg.eigvalsh(V.conj().T @ A @ V))[::-1] - ev[::-1][:m]).max() <= 1e-14 * np.sqrt(n)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("sizes", common.HEIG_BATCHES + [[(1000, 200), (997, 300), (1003, 100)]], ids=lambda s: "x".join(str(n) for n, _ in s))
def test_hermitian_eigensolver_batched(gpu_ctx, cplx, sizes):
    """The charge blocks of one bond side by side: one cooperative launch per panel step, each matrix on its own slice of the grid
    with its own arrival-counter barrier (tridiag_panel_batched_kernel)."""
    common.check_heig_batched(gpu_ctx, cplx, sizes)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,m", common.TWO_STAGE_SIZES)
def test_hermitian_eigensolver_two_stage(gpu_ctx, monkeypatch, cplx, n, m):
    """panel_qr_kernel + GEMM block updates + bulge_chase_kernel + q2_apply_kernel + compact-WY panels, forced on at small sizes."""
    common.check_heig_two_stage(gpu_ctx, cplx, n, m, monkeypatch)


@pytest.mark.parametrize("cplx,n,m", [(False, 2500, 900), (True, 4100, 300)])
def test_hermitian_eigensolver_two_stage_default_threshold(gpu_ctx, cplx, n, m):
    """Sizes that take the two-stage path by default (real n >= 2048, complex n >= 4096), against numpy.eigh."""
    rng = np.random.default_rng(n + m)
    X = _rand(rng, (n, n), cplx) * np.exp(-np.arange(n) * (14.0 / n))[None, :]
    A = X.conj().T @ X
    A /= np.trace(A).real
    w, V, it = cb.k_heig(A, m, ctx=gpu_ctx)
    ev = np.linalg.eigvalsh(A)[::-1]
    assert np.abs(w - ev).max() <= 1e-14 * np.sqrt(n)
    assert np.abs(V.conj().T @ V - np.eye(m)).max() <= 1e-13
    assert np.abs(A @ V - V @ (V.conj().T @ A @ V)).max() <= 2e-14 * np.sqrt(n)


@pytest.mark.parametrize("mode", ["1", "2", "3"])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False)])
def test_gemm_narrow_and_3m_tiles(gpu_ctx, monkeypatch, mode, ta, tb):
    """The 64-column tile and its 3-multiplication complex variant (3 DMMA per complex MAC: Re = P1 - P2, Im = P3 - P1 - P2)."""
    monkeypatch.setenv("CB200_K_GEMM_MODE", mode)
    rng = np.random.default_rng(int(mode) * 7 + ta + 2 * tb)
    for cplx in (True, False):
        for M, N, K in [(257, 131, 95), (64, 640, 33), (533, 300, 534)]:
            A = _rand(rng, (M, K), cplx)
            B = _rand(rng, (K, N), cplx)
            a = (A.T if ta else A).ravel(order="F")
            b = (B.T if tb else B).ravel(order="F")
            c, _ = cb.k_gemm(a, b, np.zeros(M * N), M, N, [K], [0], [0], K if ta else M, N if tb else K, 0, M, trans_a=ta, trans_b=tb,
                             conj_a=cplx and ta, ctx=gpu_ctx)
            want = (A.conj() if (cplx and ta) else A) @ B
            assert np.abs(c.reshape(N, M).T - want).max() <= 2e-13 * K * max(1.0, np.abs(want).max())
