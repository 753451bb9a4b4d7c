"""rho-defect measurement (second half): the operator of RhoOp / OpenRhoOp (src/chain.cpp:65-185) written as a
product of RhoDefectCells (src/f_data.cpp:94-111), its matrix elements <psi_i| rho |psi_j> on the device (cb200_mps_mpo_element)
and the full Analyze table {energy, momentum, rho} (src/dmrg.h:1118-1174).
Pins: the reference documents the expected rho eigenvalues -- (1 +- sqrt5)/2 for golden, (3 +- sqrt13)/2 and +-1 for Haagerup
(src/dmrg.h:1100-1105) -- and rho commutes with H."""
import numpy as np
import pytest

import chain_b200 as cb
from chain_b200 import driver, models
import common
from oracle import dmrg as od, ed, models as om

GOLDEN_RHO = [(1 + 5 ** 0.5) / 2, (1 - 5 ** 0.5) / 2]
HAAGERUP_RHO = [(3 + 13 ** 0.5) / 2, (3 - 13 ** 0.5) / 2, 1.0, -1.0]


def dense_rho(G, n):
    li, lo = "abcdefgh"[:n], "ijklmnop"[:n]
    expr = ",".join("%s%s%s%s" % (li[j], lo[j], li[(j + 1) % n], lo[(j + 1) % n]) for j in range(n)) + "->" + lo + li
    d = G.shape[0]
    return np.einsum(expr, *([G] * n), optimize=True).reshape(d ** n, d ** n)


@pytest.mark.parametrize("st,n,cp,want", [("golden", 6, [-1.0], GOLDEN_RHO), ("haagerup", 4, [0.0, -1.0, 0.0], HAAGERUP_RHO)])
def test_cell_product_is_the_rho_defect(st, n, cp, want):
    """Product-side and oracle-side cells agree; on the exact low-energy eigenvectors the operator has the documented eigenvalues
    and commutes with H."""
    G = models.rho_defect_cell(st)
    Go = (om.GoldenFData() if st == "golden" else om.HaagerupFData()).rho_defect_cell()
    assert np.abs(G - Go).max() <= 1e-14
    R = dense_rho(G, n)
    H = ed.build_sparse_h(st, "p", n, 2.0, cp)[0].toarray()
    w, v = np.linalg.eigh(H)
    k = 6
    Rk = v[:, :k].T @ R @ v[:, :k]
    assert np.abs(v[:, :k].T @ (H @ R - R @ H) @ v[:, :k]).max() <= 1e-12
    for x in np.linalg.eigvals(Rk).real:
        assert min(abs(x - t) for t in want) <= 1e-10


def _states(st, n, cp, q, nstates, nsweep=14):
    site, W, qk = common.build_model(st, "p", n, cp)
    done, ens = [], []
    for s in range(nstates):
        psi0 = common.warm_start(site, W, n, q, seed=3 + s)
        en, psi = od.dmrg(W, done, psi0, od.Sweeps(nsweep, [10, 20, 40, 80], [1e-6, 1e-8, 1e-10, 1e-10]), weight=1000.0)
        done.append(psi)
        ens.append(en)
    return site, ens, done


def test_rho_matrix_elements_match_oracle_golden(emu_ctx):
    site, ens, states = _states("golden", 6, [-1.0], 0, 3)
    G = models.rho_defect_cell("golden")
    mpos = [cb.MPO(m, None, None, nq=1) for m in models.rho_defect_mpos("golden", 6)]
    cbs = [common.to_cb(p) for p in states]
    R_o = np.array([[od.rho_element(a, G, b) for b in states] for a in states])
    R_g = np.array([[sum(cb.mpo_element(a, O, b, ctx=emu_ctx) for O in mpos) for b in cbs] for a in cbs])
    assert np.abs(R_o - R_g).max() <= 1e-11
    for x in np.linalg.eigvals(R_g).real:                        # converged eigenstates of H: rho is diagonal-able on them
        assert min(abs(x - t) for t in GOLDEN_RHO) <= 1e-5
    # open-chain operator (OpenRhoOp): no closure factor
    mo = [cb.MPO(m, None, None, nq=1) for m in models.rho_defect_mpos("golden", 6, periodic=False)]
    assert len(mo) == 1
    assert abs(cb.mpo_element(cbs[0], mo[0], cbs[1], ctx=emu_ctx) - od.rho_element(states[0], G, states[1], periodic=False)) <= 1e-11


def test_rho_through_the_z3_fourier_map(emu_ctx):
    # haagerupq states are first taken to the haagerup basis (removeQNs + Z3FourierTransform, src/dmrg.h:781-784)
    site, ens, states = _states("haagerupq", 4, [0.0, -1.0, 0.0], 0, 1, nsweep=10)
    U = models.z3_fourier_matrix()
    assert np.abs(U - om.z3_fourier_matrix()).max() <= 1e-15 and np.abs(U.conj().T @ U - np.eye(6)).max() <= 1e-14
    psi_h = models.to_haagerup_basis(common.to_cb(states[0]))
    G = models.rho_defect_cell("haagerupq")
    r = sum(cb.mpo_element(psi_h, cb.MPO(m, None, None, nq=1), psi_h, ctx=emu_ctx) for m in models.rho_defect_mpos("haagerupq", 4))
    ho = od.MPS([np.einsum("lsr,st->ltr", a.astype(complex), U) for a in states[0].A], [np.zeros(len(q), int) for q in states[0].q], np.zeros(6, int), 1)
    assert abs(r - od.rho_element(ho, G, ho)) <= 1e-11
    assert abs(r.imag) <= 1e-9 and min(abs(r.real - t) for t in HAAGERUP_RHO) <= 1e-4


def test_full_analysis_golden(emu_ctx, tmp_path):
    sim = driver.Simulation("golden", "p", 6, 100, 1e-8, 1e-9, "-1", 2.0, 0, 3, out=str(tmp_path), seed=1, ctx=emu_ctx, log=lambda *a: None)
    sim.run()
    rows = driver.analyze(sim, log=lambda *a: None)
    want_e = common.ed_level("golden", "p", 6, [-1.0], None)[:3]
    assert np.allclose([r[0] for r in rows], want_e, atol=2e-9)
    assert [r[1] for r in rows] == [0.0, 3.0, 0.0]
    for r in rows:
        assert len(r) == 3 and min(abs(r[2] - t) for t in GOLDEN_RHO) <= 2e-5    # rounded to 1e-6 like the reference
    assert abs(rows[0][2] - GOLDEN_RHO[0]) <= 2e-5                               # ED: the ground state carries (1+sqrt5)/2
    m = open(sim.m_path).read()
    assert "\nrho = {\n" in m
    an = open(sim.an_path).read().splitlines()
    assert an[1].rstrip(",").count(",") == 2
