"""Pins the oracle AND the product-side model tables to data computed by the REFERENCE'S OWN CODE.

tests/golden/reference_models.json is the output of /root/reference/src/{golden_site,haagerup_site,haagerup_q_site,f_data}.cpp compiled
unmodified against tests/itensor_stub (tests/golden/from_reference.py; oracle/Makefile).  Here:
  * when /root/reference is present the dump is regenerated and must equal the committed fixture byte for byte;
  * every site operator, state label, Z3 sector charge, F-symbol and RhoDefectCell entry of oracle/models.py and of
    chain_b200/models.py equals the reference's;
  * the Hamiltonian term lists equal the reference's `mpo += ...` lists term by term (all five boundary conditions);
  * the compressed MPOs of chain_b200/models.py and oracle/mpo.py, contracted densely, equal H assembled from the reference's terms;
  * the ED known answers (tests/golden/ed_levels.json, which pin the DMRG restatement) equal the levels of H assembled from nothing but
    the reference dump (tests/golden/ed_levels_reference.json).
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import from_reference as fr  # noqa: E402
from oracle import models as omodels, mpo as ompo  # noqa: E402
from chain_b200 import models as pmodels  # noqa: E402
from common import ed_level  # noqa: E402

DUMP = fr.load_fixture()
SITES = ["golden", "haagerup", "haagerupq"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference sources are not on this machine")
def test_fixture_is_what_the_reference_code_prints_today():
    assert fr.run_reference() == open(fr.FIXTURE).read()


@pytest.mark.parametrize("site", SITES)
def test_site_operators_equal_the_reference(site):
    osite, psite = omodels.make_site(site), pmodels.make_siteset(site, 6)
    names = list(DUMP[site]["ops"])
    assert len(names) >= {"golden": 4, "haagerup": 28, "haagerupq": 37}[site]
    for nm in names:
        ref = fr.ref_operator(DUMP, site, nm)
        if site == "haagerupq" and nm == "id":
            # the reference's HaagerupQSite::id() is 1/3 on the diagonal (src/haagerup_q_site.cpp:41-48); the term lists scale with it
            assert np.allclose(ref, np.eye(6) / 3, atol=0)
        assert np.abs(osite.ops[nm] - ref).max() <= 1e-16, (site, nm)
        assert np.abs(psite.op(nm) - ref).max() <= 1e-16, (site, nm)
    assert DUMP[site]["d"] == osite.d == psite.d


def test_states_and_sectors_equal_the_reference():
    for site in SITES:
        osite, psite = omodels.make_site(site), pmodels.make_siteset(site, 6)
        for label, val in DUMP[site]["states"].items():
            assert osite.state(label) == val - 1 and psite.state(label) == val - 1, (site, label)
    q = DUMP["haagerupq"]
    assert q["sector_charges"] == list(omodels.make_site("haagerupq").charges) == list(pmodels.make_siteset("haagerupq", 6).charges)
    assert q["qn_mod"] == 3 and q["arrow"] == 1


@pytest.mark.parametrize("cat", ["golden", "haagerup"])
def test_fsymbols_and_rho_defect_cell_equal_the_reference(cat):
    d = DUMP[cat + "_fdata"]
    of = omodels.GoldenFData() if cat == "golden" else omodels.HaagerupFData()
    pf = pmodels.golden_fdata() if cat == "golden" else pmodels.haagerup_fdata()
    rk = d["rk"]
    ref = np.zeros((rk,) * 6)
    for i, j, k, l, m, n, v in d["fsymbols"]:
        ref[i - 1, j - 1, k - 1, l - 1, m - 1, n - 1] = v
    it = np.ndindex(*ref.shape)
    for idx in it:
        a = tuple(x + 1 for x in idx)
        assert abs(of.fsymbol(*a) - ref[idx]) <= 1e-16 and abs(pf.F(*a) - ref[idx]) <= 1e-16, (cat, a)
    G = np.zeros((rk,) * 4)
    for i1, j1, i2, j2, v in d["rho_defect_cell"]:
        G[i1 - 1, j1 - 1, i2 - 1, j2 - 1] = v
    assert np.abs(of.rho_defect_cell() - G).max() <= 1e-16
    assert np.abs(pmodels.rho_defect_cell(cat) - G).max() <= 1e-16


def _canon(terms):
    """term list -> {((site, op), ...) sorted: summed coefficient}: AutoMPO sums repeated operator strings."""
    acc = {}
    for c, ops in terms:
        key = tuple(sorted((int(s), o) for o, s in ops))
        acc[key] = acc.get(key, 0) + complex(c)
    return acc


HCASES = [(s, h) for s in SITES for h in DUMP[s]["hamiltonians"]]


@pytest.mark.parametrize("site,h", HCASES, ids=lambda x: x if isinstance(x, str) else "%s-n%d-%s" % (x["bc"], x["n"], x["couplings"]))
def test_hamiltonian_terms_equal_the_reference(site, h):
    ref = _canon([(complex(*t["c"]), t["ops"]) for t in h["terms"]])
    nterms = len(h["terms"])
    got_o = omodels.hamiltonian_terms(site, h["bc"], h["n"], h["U"], h["couplings"])
    got_p = pmodels.make_siteset(site, h["n"]).terms(h["bc"], h["n"], h["U"], list(h["couplings"]))
    assert len(got_o) == nterms and len(got_p) == nterms
    for got in (_canon(got_o), _canon(got_p)):
        assert set(got) == set(ref)
        for k in ref:
            assert abs(got[k] - ref[k]) <= 4e-16 * max(1.0, abs(ref[k])), (site, h["bc"], k, got[k], ref[k])


@pytest.mark.parametrize("site,bc", [(s, bc) for s in SITES for bc in ("p", "o", "s", "d", "sp")])
def test_compressed_mpos_equal_h_of_the_reference_terms(site, bc):
    n, cp = (6, [-1.0]) if site == "golden" else (4, [-0.5, 0.25, 1.5])
    H = fr.ref_sparse_h(DUMP, site, bc, n, cp).toarray()
    ss = pmodels.make_siteset(site, n)
    Hp = pmodels.mpo_to_dense(ss.hamiltonian(bc, n, 2.0, cp))
    assert np.abs(Hp - H).max() <= 1e-12 * max(1.0, np.abs(H).max())
    osite = omodels.make_site(site)
    W = ompo.build_mpo(omodels.hamiltonian_terms(site, bc, n, 2.0, cp), osite, n)
    Ho = ompo.mpo_to_dense(W[0] if isinstance(W, tuple) else W)
    assert np.abs(Ho - H).max() <= 1e-12 * max(1.0, np.abs(H).max())


def test_ed_known_answers_follow_from_the_reference_dump():
    with open(fr.ED_FIXTURE) as f:
        ref = json.load(f)
    assert len(ref["levels"]) >= 11
    for c in ref["levels"]:
        want = np.array(c["levels"])
        got = np.array(ed_level(c["site"], c["bc"], c["n"], c["couplings"], c["charge"]))
        k = min(len(want), len(got))
        # compared as sorted level lists; Lanczos may report a different multiplicity inside an exactly degenerate multiplet
        uw, ug = np.unique(np.round(want[:k], 8)), np.unique(np.round(got[:k], 8))
        m = min(len(uw), len(ug))
        assert m >= 1 and np.allclose(uw[:m], ug[:m], atol=1e-8), c
        assert abs(want[0] - got[0]) <= 1e-10


def test_ed_from_reference_dump_recomputed_live():
    # cheap live rediagonalisation from the dump alone (golden chain; the d = 6 cases are in ed_levels_reference.json)
    w = fr.ref_levels(DUMP, "golden", "p", 6, [-1.0], None, 5)
    assert np.allclose(w, ed_level("golden", "p", 6, [-1.0], None)[:5], atol=1e-12)
    assert np.allclose(w, [-4.698543184702, -4.548658557883, -4.311825964952, -3.0, -2.906490107234], atol=2e-12)
