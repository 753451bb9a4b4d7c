"""Product-side Hamiltonians and compressed MPOs (chain_b200/models.py: Golden / Haagerup / HaagerupQ ::Hamiltonian of
src/golden_site.cpp:57-126, src/haagerup_site.cpp:111-238, src/haagerup_q_site.cpp:355-536, toMPO-compressed) against the oracle's
exact diagonalisation, which assembles H as a sparse matrix straight from the term lists and never sees an MPO.  The spectrum is read
off the MPO the way the reference's own cross-check does (mode 7, Test(), src/dmrg.h:1283-1300: driver.exact_levels)."""
import numpy as np
import pytest

from chain_b200 import driver, models
from oracle import ed


CASES = [
    # site, bc, n, couplings, U
    ("golden", "p", 6, [-1.0], 2.0),
    ("golden", "o", 7, [-1.0], 2.0),
    ("golden", "s", 6, [-1.0], 2.0),
    ("golden", "sp", 6, [-1.0], 2.0),
    ("golden", "d", 6, [-1.0], 2.0),
    ("golden", "p", 5, [0.7], 3.5),
    ("haagerup", "p", 4, [0.0, -1.0, 0.0], 2.0),
    ("haagerup", "o", 4, [-1.0, 0.0, 0.0], 2.0),
    ("haagerup", "s", 4, [0.3, -1.0, 0.2], 2.0),
    ("haagerupq", "p", 4, [0.0, -1.0, 0.0], 2.0),      # complex128 MPO
    ("haagerupq", "p", 4, [-1.0, 0.0, 0.0], 2.0),
    ("haagerupq", "o", 4, [0.0, 0.0, -1.0], 2.0),
    ("haagerupq", "s", 4, [0.4, -0.6, 0.5], 1.0),
]


@pytest.mark.parametrize("st,bc,n,cp,U", CASES, ids=lambda v: str(v))
def test_mpo_spectrum_matches_independent_ed(st, bc, n, cp, U):
    try:
        want = ed.lowest_levels(st, bc, n, U, cp, k=8)
    except (ValueError, KeyError) as err:
        pytest.skip("oracle has no such boundary condition: %s" % err)
    ss = models.make_siteset(st, n)
    mpo = ss.hamiltonian(bc, n, U, cp)
    got = driver.exact_levels(mpo.sites, 8)
    assert np.allclose(got, np.sort(want)[:8], atol=1e-9)
    # every MPO tensor respects the Z_N flux rule the C ABI derives its block structure from (include/chain_b200.h)
    sq = np.asarray(ss.charges)
    for W, ql, qr in zip(mpo.sites, mpo.link_charges[:-1], mpo.link_charges[1:]):
        flux = (np.asarray(ql)[:, None, None, None] + sq[None, :, None, None] - sq[None, None, :, None] - np.asarray(qr)[None, None, None, :]) % ss.nq
        assert np.abs(np.asarray(W)[flux != 0]).max(initial=0.0) == 0.0


def test_compressed_mpo_bond_dimensions():
    # SURVEY App. D (what ITensor's SVD-compressing toMPO yields, bond by bond)
    k = lambda st, bc, n, cp: [len(q) for q in models.make_siteset(st, n).hamiltonian(bc, n, 2.0, cp).link_charges][1:-1]
    assert k("golden", "p", 8, [-1.0]) == [5, 10, 10, 10, 10, 10, 5]
    assert k("golden", "o", 8, [-1.0]) == [4, 6, 6, 6, 6, 6, 4]
    assert k("golden", "s", 8, [-1.0]) == [4, 6, 6, 6, 6, 6, 4]
    assert k("haagerup", "o", 8, [0.0, -1.0, 0.0]) == [8, 14, 14, 14, 14, 14, 8]
    assert k("haagerup", "p", 8, [0.0, -1.0, 0.0]) == [17, 23, 26, 26, 26, 23, 17]
    assert k("haagerupq", "p", 8, [-1.0, 0.0, 0.0]) == [11, 26, 26, 26, 26, 26, 11]
    hq = models.make_siteset("haagerupq", 8).hamiltonian("p", 8, 2.0, [0.0, -1.0, 0.0])
    mid = np.asarray(hq.link_charges[4])
    assert len(mid) == 26 and np.bincount(mid, minlength=3).tolist() == [10, 8, 8]
    ho = models.make_siteset("haagerupq", 8).hamiltonian("o", 8, 2.0, [0.0, -1.0, 0.0])
    assert np.bincount(np.asarray(ho.link_charges[4]), minlength=3).tolist() == [6, 4, 4]      # 14 = 2 + 4 + 4 + 4
