"""Mode 1/5 momentum assignment (the translation half): one-site translation by L-1 swap gates with truncated
SVDs (Swap, src/chain.cpp:486-507; Measure("translation"), src/dmrg.h:855-858), T_ij = innerC(psi_i, T psi_j) (:1063) and
AnalyzeWithoutRho (:1177-1218).  Host logic on the emulation backend against the oracle restatement; the known answer is that
translation eigenvalues of a periodic chain are L-th roots of unity (src/dmrg.h:661-666)."""
import numpy as np
import pytest

import chain_b200 as cb
from chain_b200 import driver
import common
from oracle import dmrg as od


def _states(st, bc, n, cp, q, nstates, nsweep=12):
    site, W, qk = common.build_model(st, bc, n, cp)
    done, ens = [], []
    for s in range(nstates):
        psi0 = common.warm_start(site, W, n, q, seed=3 + s)
        en, psi = od.dmrg(W, done, psi0, od.Sweeps(nsweep, [10, 20, 40, 80], [1e-6, 1e-8, 1e-10, 1e-10]), weight=1000.0)
        done.append(psi)
        ens.append(en)
    return site, ens, done


@pytest.mark.parametrize("st,n,cp,q", [("golden", 6, [-1.0], 0), ("haagerupq", 6, [-1.0, 0.0, 0.0], 1)])
def test_translate_and_overlap_match_oracle(emu_ctx, st, n, cp, q):
    site, ens, states = _states(st, "p", n, cp, q, 2)
    for psi in states:
        to = od.swap_translate(psi, 1e-5)
        tg = cb.translate(common.to_cb(psi), site.charges, 1e-5, ctx=emu_ctx)
        assert tg.bond_dims() == to.bond_dims()
        tgo = common.to_oracle(tg, site)
        for a in states:
            zo = od.overlap(a, to)
            zg = cb.overlap(common.to_cb(a), tg, site.charges, ctx=emu_ctx)
            assert abs(zo - zg) <= 1e-10
            assert abs(od.overlap(a, tgo) - zg) <= 1e-11          # the overlap kernel itself, on identical inputs
    # T psi is (up to the 1e-5 cutoff) psi times an L-th root of unity for a non-degenerate level
    z = cb.overlap(common.to_cb(states[0]), cb.translate(common.to_cb(states[0]), site.charges, 1e-5, ctx=emu_ctx), site.charges, ctx=emu_ctx)
    if st == "golden":
        assert abs(abs(z) - 1) <= 1e-4
        k = np.angle(z) / (2 * np.pi) * n
        assert abs(k - round(k)) <= 1e-3


def test_momentum_assignment_golden(emu_ctx, tmp_path):
    # three lowest levels of the golden chain L=6: ED momenta 0, +-? -- compare device path with the oracle's AnalyzeWithoutRho
    site, ens, states = _states("golden", "p", 6, [-1.0], 0, 3, nsweep=16)
    T_o = np.array([[od.overlap(a, od.swap_translate(b, 1e-5)) for b in states] for a in states])
    tg = [cb.translate(common.to_cb(b), site.charges, 1e-5, ctx=emu_ctx) for b in states]
    T_g = np.array([[cb.overlap(common.to_cb(a), t, site.charges, ctx=emu_ctx) for t in tg] for a in states])
    assert np.abs(T_o - T_g).max() <= 1e-10
    want = od.momentum_analysis(ens, T_o, 6)
    got = driver.analyze_without_rho(ens, T_g, 6)
    assert [m for _, m in got] == [m for _, m in want]
    assert np.allclose([e for e, _ in got], [e for e, _ in want], atol=1e-10)
    for _, m in got:
        assert abs(m - round(m)) <= 1e-3                          # integer momenta (L-th roots of unity)
    p = tmp_path / "x.an"
    driver.dump_matrix(str(p), [list(r) for r in got])
    txt = open(p).read()
    assert txt.startswith("{\n{") and txt.endswith("}\n}\n") and txt.count("\n") == len(got) + 2
    pm = tmp_path / "x.m"
    driver.dump_mathematica(str(pm), "translation", T_g)
    assert open(pm).read().startswith("translation = {\n{") and " + I * " in open(pm).read()


def test_driver_mode5_writes_momentum_table(emu_ctx, tmp_path):
    sim = driver.Simulation("golden", "p", 6, 100, 1e-8, 1e-9, "-1", 2.0, 0, 3, out=str(tmp_path), seed=1, ctx=emu_ctx, log=lambda *a: None)
    sim.run()
    rows = driver.analyze(sim, log=lambda *a: None)
    want = common.ed_level("golden", "p", 6, [-1.0], None)[:3]
    assert np.allclose([r[0] for r in rows], want, atol=2e-9)
    assert [r[1] for r in rows] == [0.0, 3.0, 0.0]               # golden chain L=6: the second level sits at momentum pi
    assert open(sim.en_path).read().splitlines() == ["{%d,%.10f}," % (i + 1, e) for i, e in enumerate(sim.energies)]
    m = open(sim.m_path).read()
    assert m.startswith("energy = {\n") and "\ntranslation = {\n" in m
    an = open(sim.an_path).read().splitlines()
    assert an[0] == "{" and an[-1] == "}" and len(an) == 5
