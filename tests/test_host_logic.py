"""Host-logic tests (CPU): the engine's block bookkeeping / task-list generation / sequencing, executed on the
serial emulation backend (tests/hostemu) and compared with the oracle.  These do NOT test the CUDA kernels; the
-m gpu tests run the same bodies through the product library."""
import numpy as np
import pytest

import chain_b200 as cb
import common


@pytest.mark.parametrize("case", common.BOND_CASES, ids=lambda c: "%s-%s-%d-q%d" % (c[0], c[1], c[2], c[4]))
def test_bond_ops_match_oracle(emu_ctx, case):
    common.check_bond_ops(emu_ctx, case)


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("case", common.BOND_CASES, ids=lambda c: "%s-%s-%d-q%d" % (c[0], c[1], c[2], c[4]))
def test_bond_apply_block_packed_boundary(emu_ctx, case, shuffle):
    common.check_bond_blocks(emu_ctx, case, shuffle)


@pytest.mark.parametrize("case", common.DMRG_CASES, ids=lambda c: "%s-%s-%d-q%d" % (c[0], c[1], c[2], c[4]))
def test_dmrg_trajectory_matches_oracle(emu_ctx, case):
    common.check_dmrg_trajectory(emu_ctx, case, nsweep=8)


def test_task_list_gemm_segments_and_split(emu_ctx):
    rng = np.random.default_rng(0)
    M, N = 37, 29
    Ks = [5, 0, 13]
    A = rng.standard_normal(M * 40)
    B = rng.standard_normal(40 * N)
    a_off, b_off = [0, 0, M * 7], [0, 0, 9]
    c, _ = cb.k_gemm(A, B, np.zeros(M * N), M, N, Ks, a_off, b_off, M, 40, 0, M, ctx=emu_ctx)
    want = sum(A[ao:ao + M * k].reshape(k, M).T @ np.stack([B[bo + j * 40: bo + j * 40 + k] for j in range(N)], axis=1)
               for ao, bo, k in zip(a_off, b_off, Ks) if k)
    assert np.abs(c.reshape(N, M).T - want).max() < 1e-12


@pytest.mark.parametrize("rows,cols", [(90, 70), (30, 70), (10, 100), (64, 64), (5, 3), (1, 1)])
@pytest.mark.parametrize("cplx", [False, True])
def test_block_jacobi_svd_driver(emu_ctx, rows, cols, cplx):
    rng = np.random.default_rng(rows * 100 + cols)
    X = rng.standard_normal((rows, cols)) + (1j * rng.standard_normal((rows, cols)) if cplx else 0)
    if cols > 5:
        X[:, 5] = 0
    xv, v, w, sweeps = cb.k_svd(X, ctx=emu_ctx)
    s = np.linalg.svd(X, compute_uv=False)
    k = min(rows, cols)
    assert np.abs(w[:k] - s ** 2).max() < 1e-11 * max(1.0, s[0] ** 2)
    assert np.abs(v.conj().T @ v - np.eye(cols)).max() < 1e-12
    assert np.abs(X @ v - xv).max() < 1e-12 * max(1.0, s[0])


def test_ground_state_kat_golden(emu_ctx):
    common.check_ground_state_kat(emu_ctx, "golden", "p", 6, [-1.0], 0, tol=2e-9)


def test_excited_state_in_charge_sector(emu_ctx):
    common.check_excited_qn(emu_ctx)


def test_excited_state_in_qn_sector_converged(emu_ctx):
    """north_star tolerances (energies 1e-10 relative, EE 1e-8) on the penalised Z3 state, at convergence (100 sweeps, ~25 s)."""
    common.check_excited_qn_converged(emu_ctx)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,m,decay", [(1, 1, 0.0), (2, 1, 0.5), (3, 3, 0.3), (37, 20, 0.4), (90, 90, 0.0), (120, 40, 0.1)])
def test_hermitian_eigensolver_driver(emu_ctx, cplx, n, m, decay):
    rng = np.random.default_rng(n * 10 + m + cplx)
    X = (rng.standard_normal((n + 5, n)) + (1j * rng.standard_normal((n + 5, n)) if cplx else 0)) * (10.0 ** (-decay * np.arange(n)))
    if n > 30:
        X[:, 7] = X[:, 3]
    A = X.conj().T @ X
    A /= np.trace(A).real
    w, V, it = cb.k_heig(A, m, ctx=emu_ctx)
    ev = np.linalg.eigvalsh(A)[::-1]
    assert np.abs(w - ev).max() <= 1e-14 * np.sqrt(n)
    assert np.abs(V.conj().T @ V - np.eye(m)).max() <= 1e-13
    assert np.abs(A @ V - V @ (V.conj().T @ A @ V)).max() <= 2e-14 * np.sqrt(n)
    assert np.abs(np.sort(np.linalg.eigvalsh(V.conj().T @ A @ V))[::-1] - ev[:m]).max() <= 1e-14 * np.sqrt(n)


@pytest.mark.parametrize("case", [common.DMRG_CASES[0], common.DMRG_CASES[3], common.DMRG_CASES[4]], ids=["golden", "haagerupq-real", "haagerupq-complex"])
def test_mps_boundary_device_packing_equals_host_loop(emu_ctx, case, monkeypatch):
    common.check_mps_boundary_paths(emu_ctx, case, monkeypatch)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("sizes", common.HEIG_BATCHES, ids=lambda s: "x".join(str(n) for n, _ in s))
def test_hermitian_eigensolver_batched_driver(emu_ctx, cplx, sizes):
    common.check_heig_batched(emu_ctx, cplx, sizes)


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("case", [common.DMRG_CASES[0], common.DMRG_CASES[3], common.DMRG_CASES[4]], ids=["golden", "haagerupq-real", "haagerupq-complex"])
def test_block_packed_mps_boundary(emu_ctx, case, shuffle):
    common.check_packed_mps_boundary(emu_ctx, case, shuffle)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,m", common.TWO_STAGE_SIZES[:2])
def test_hermitian_eigensolver_two_stage_driver(emu_ctx, monkeypatch, cplx, n, m):
    """Host logic of the two-stage path (panel/GEMM/bulge-chase/back-transformation plumbing) on the emulation backend."""
    common.check_heig_two_stage(emu_ctx, cplx, n, m, monkeypatch)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("direction", ["left", "right"])
def test_true_svd_branch_below_cutoff_1e12(emu_ctx, cplx, direction):
    common.check_true_svd_branch(emu_ctx, cplx, direction, D=48)


@pytest.mark.parametrize("case", [common.BOND_CASES[2], common.BOND_CASES[4]], ids=["dense-real", "z3-complex"])
def test_truncation_falls_back_to_jacobi_when_orthonormalisation_fails(emu_ctx, monkeypatch, case):
    common.check_truncation_survives_ns_failure(emu_ctx, case, monkeypatch)


def test_dmrg_run_on_the_jacobi_fallback_matches_oracle(emu_ctx, monkeypatch):
    """A whole run in which EVERY truncation takes the fallback: energy still 1e-10 from the oracle's."""
    from oracle import dmrg as od
    site, W, qk = common.build_model("haagerupq", "p", 6, [-1.0, 0.0, 0.0])
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    psi0 = od.product_mps(site, 6, 0, np.random.default_rng(0))
    sw = ([10, 20, 40], [1e-6, 1e-7, 1e-8])
    eo, _ = od.dmrg(W, [], psi0, od.Sweeps(4, *sw))
    monkeypatch.setenv("CB200_NS_MAXIT", "-1")
    eg, _ = cb.dmrg(H, [], common.to_cb(psi0), cb.Sweeps(4, *sw), ctx=emu_ctx)
    assert abs(eo - eg) <= 1e-10 * abs(eo), (eo, eg)
