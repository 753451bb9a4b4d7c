"""Tolerances of the shared parity bodies against a DIFFERENT but equally valid arithmetic (tools/rounding_audit.py, short form).
The host emulation perturbs every GEMM result the way another summation order / the 3M complex product would (tests/hostemu/backend_emu.cpp,
CB200_EMU_ROUNDING); the GPU kernels differ from the serial loops in exactly that, so a body that only passes with the unperturbed loops
would be a GPU test that passes by luck.  Run here at 8x the modelled noise; the long form (whole host suite at 1x / 8x / 64x, the converged
penalised state per seed) is run by hand and recorded in DESIGN.md section 6."""
import pytest

import common


@pytest.fixture
def noisy(monkeypatch):
    monkeypatch.setenv("CB200_EMU_ROUNDING", "11")
    monkeypatch.setenv("CB200_EMU_ROUNDING_ULPS", "8")


@pytest.mark.parametrize("case", [common.BOND_CASES[2], common.BOND_CASES[4]], ids=["dense-real", "z3-complex"])
def test_bond_ops_under_perturbed_rounding(emu_ctx, noisy, case):
    common.check_bond_ops(emu_ctx, case)


@pytest.mark.parametrize("case", [common.DMRG_CASES[0], common.DMRG_CASES[4]], ids=["golden", "haagerupq-complex"])
def test_dmrg_trajectory_under_perturbed_rounding(emu_ctx, noisy, case):
    common.check_dmrg_trajectory(emu_ctx, case, nsweep=8)


def test_perturbation_is_active(emu_ctx, monkeypatch):
    """The knob really changes the bits (otherwise the tests above prove nothing) and is deterministic per call."""
    import numpy as np
    import chain_b200 as cb
    rng = np.random.default_rng(0)
    M, N, K = 8, 5, 33
    A, B = rng.standard_normal(M * K), rng.standard_normal(K * N)
    run = lambda: cb.k_gemm(A, B, np.zeros(M * N), M, N, [K], [0], [0], M, K, 0, M, ctx=emu_ctx)[0]
    monkeypatch.delenv("CB200_EMU_ROUNDING", raising=False)
    c0 = run()
    monkeypatch.setenv("CB200_EMU_ROUNDING", "3")
    c1, c2 = run(), run()
    assert np.array_equal(c1, c2) and not np.array_equal(c0, c1)
    assert np.abs(c1 - c0).max() <= 4e-16 * np.abs(A).max() * np.abs(B).max() * K
