"""Shared bodies of the parity tests: the same checks run against the host-emulation build (CPU, host logic only)
and against the product library on a B200 (-m gpu).  The oracle (oracle/) is the checker."""
import json
import os

import numpy as np

import chain_b200 as cb
from oracle import models, mpo as ompo, dmrg as od

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PRODUCT_LIB = os.path.join(ROOT, "chain_b200", "libchain_b200.so")


def golden():
    with open(os.path.join(HERE, "golden", "ed_levels.json")) as f:
        return json.load(f)


def ed_level(site, bc, n, couplings, charge):
    for c in golden()["levels"]:
        if c["site"] == site and c["bc"] == bc and c["n"] == n and c["couplings"] == list(couplings) and c["charge"] == charge:
            return c["levels"]
    raise KeyError((site, bc, n, couplings, charge))


def build_model(st, bc, n, couplings, U=2.0):
    site = models.make_site(st)
    W, qk = ompo.build_mpo(models.hamiltonian_terms(st, bc, n, U, couplings), site, n)
    return site, W, qk


def to_cb(psi):
    return cb.MPS(psi.A, psi.q, nq=psi.nq, center=psi.center or 0)


def to_oracle(m, site):
    return od.MPS([np.array(a) for a in m.sites], [np.array(q, dtype=np.int64) for q in m.link_charges], site.charges, site.nq, center=m.center)


def warm_start(site, W, n, charge, seed, maxdim=8, nsweep=2):
    """A non-trivial common start state: a couple of oracle sweeps from the reference's product state.
    (For QN sites the product state itself triggers Davidson's random restart, which no two implementations share.)"""
    rng = np.random.default_rng(seed)
    psi0 = od.product_mps(site, n, charge, rng, dtype=complex if np.iscomplexobj(W[1]) else float)
    _, psi = od.dmrg(W, [], psi0, od.Sweeps(nsweep, maxdim, 1e-10))
    return psi


def capture_bond(st, bc, n, couplings, charge, bond, seed=3, nsw=3, maxdims=(10, 20, 40)):
    """Run the oracle a few sweeps and capture L, W, R, phi of one bond update (realistic block structure)."""
    site, W, qk = build_model(st, bc, n, couplings)
    rng = np.random.default_rng(seed)
    psi0 = od.product_mps(site, n, charge, rng, dtype=complex if np.iscomplexobj(W[1]) else float)
    caps = []
    od.dmrg(W, [], psi0, od.Sweeps(nsw, list(maxdims), 1e-10), matvec_hook=caps.append)
    c = [x for x in caps if x["bond"] == bond][-1]
    return site, W, qk, c


def make_bond(ctx, site, qk, c):
    b = c["bond"]
    return cb.Bond(c["L"], c["W1"], c["W2"], c["R"], ql=c["ql"], qr=c["qr"], qkl=qk[b - 1], qkm=qk[b], qkr=qk[b + 1],
                   site_charges=site.charges, nq=site.nq, ctx=ctx)


BOND_CASES = [
    ("golden", "p", 6, [-1.0], 0, 3),
    ("golden", "s", 8, [-1.0], 0, 4),
    ("haagerup", "p", 6, [0.0, -1.0, 0.0], 0, 3),
    ("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 1, 3),
    ("haagerupq", "p", 6, [0.0, -1.0, 0.3], 2, 3),     # complex
    ("haagerupq", "o", 6, [0.0, 0.0, -1.0], 0, 2),     # complex, open chain, off-centre bond
]


def check_bond_ops(ctx, case, rtol=1e-12):
    st, bc, n, cp, q, bond = case
    site, W, qk, c = capture_bond(st, bc, n, cp, q, bond)
    B = make_bond(ctx, site, qk, c)
    phi = c["phi"]
    # LocalMPO::product
    y = B.apply(phi)
    yo = od.heff_apply(c["L"], c["W1"], c["W2"], c["R"], phi)
    assert np.abs(y - yo).max() <= rtol * np.abs(yo).max()
    # Hermiticity of H_eff on random allowed vectors (size-independent property)
    rng = np.random.default_rng(11)
    mask = (np.abs(phi) > 0) | (np.abs(yo) > 0)
    x1 = rng.standard_normal(phi.shape) * mask
    x2 = rng.standard_normal(phi.shape) * mask
    h12 = np.vdot(x1, B.apply(x2))
    h21 = np.vdot(x2, B.apply(x1))
    assert abs(h12 - np.conj(h21)) <= 1e-11 * max(1.0, abs(h12))
    # davidson
    mv = lambda v: od.heff_apply(c["L"], c["W1"], c["W2"], c["R"], v)
    evo, xo, nmvo = od.davidson(mv, phi.astype(yo.dtype), 2)
    ev, x, nmv = B.davidson(phi, 2)
    assert nmv == nmvo
    assert abs(ev - evo) <= 1e-12 * abs(evo)
    assert abs(abs(np.vdot(xo, x)) - 1) <= 1e-10
    # penalty projector (LocalMPO_MPS): y += w <o|x> o
    o = rng.standard_normal(phi.shape) * mask
    o = o / np.linalg.norm(o)
    B.set_penalty([o], 1000.0)
    yp = B.apply(phi)
    assert np.abs(yp - (yo + 1000.0 * np.vdot(o, phi) * o)).max() <= 1e-11 * max(1.0, np.abs(yp).max())
    B.set_penalty([], 0.0)
    # svdBond + makeL/makeR
    for d in ("left", "right"):
        A, Bt, qm, spec, terr = B.truncate(xo, d, 40, 1e-8)
        A1, A2, qmo, speco, terro = od.split_bond(xo, c["ql"], c["qr"], site.charges, site.nq, d, 40, 1, 1e-8)
        assert A.shape[2] == A1.shape[2]
        assert np.array_equal(np.bincount(qm, minlength=site.nq), np.bincount(qmo, minlength=site.nq))
        assert abs(terr - terro) <= 1e-14
        assert np.abs(np.sort(spec) - np.sort(speco)).max() <= 1e-12
        iso = A if d == "left" else Bt
        m = iso.shape[2] if d == "left" else iso.shape[0]
        g = np.tensordot(iso.conj(), iso, axes=([0, 1], [0, 1])) if d == "left" else np.tensordot(iso, iso.conj(), axes=([1, 2], [1, 2]))
        assert np.abs(g - np.eye(m)).max() <= 1e-12
        rec = np.tensordot(A, Bt, axes=([2], [0]))
        reco = np.tensordot(A1, A2, axes=([2], [0]))
        assert np.abs(rec - reco).max() <= 1e-9     # eigh(rho) resolves small eigenvectors only to ~eps/gap
        T = A if d == "left" else Bt
        E = B.env_update(d, T, qm)
        Eo = od.update_left(c["L"], T, c["W1"]) if d == "left" else od.update_right(c["R"], T, c["W2"])
        assert np.abs(E - Eo).max() <= rtol * max(1.0, np.abs(Eo).max())
    B.close()


def check_bond_blocks(ctx, case, shuffle=False):
    """cb200_bond_apply_blocks (ITensor QDense storage at the boundary) == block-packing of the dense product; also checks the
    block list against an independent enumeration.  shuffle=True interleaves the link charge labels (many short sectors)."""
    st, bc, n, cp, q, bond = case
    site, W, qk, c = capture_bond(st, bc, n, cp, q, bond)
    c = dict(c)
    if shuffle:
        rng = np.random.default_rng(5)
        pl, pr = rng.permutation(len(c["ql"])), rng.permutation(len(c["qr"]))
        c["ql"], c["qr"] = np.asarray(c["ql"])[pl], np.asarray(c["qr"])[pr]
        c["L"] = c["L"][pl][:, :, pl]
        c["R"] = c["R"][pr][:, :, pr]
        c["phi"] = c["phi"][pl][:, :, :, pr]
    B = make_bond(ctx, site, qk, c)
    phi = c["phi"]
    yo = od.heff_apply(c["L"], c["W1"], c["W2"], c["R"], phi)
    # the packed layout holds every allowed element exactly once
    mask = (np.asarray(c["ql"])[:, None, None, None] + np.asarray(site.charges)[None, :, None, None]
            + np.asarray(site.charges)[None, None, :, None] - np.asarray(c["qr"])[None, None, None, :]) % site.nq == 0
    assert B.phi_blocks_elems == int(mask.sum())
    blocks = B.phi_blocks()
    assert sum(int(np.prod(d)) for _, d in blocks) == B.phi_blocks_elems
    assert [o for o, _ in blocks] == list(np.cumsum([0] + [int(np.prod(d)) for _, d in blocks[:-1]]))
    xb = B.pack(phi)
    assert xb.size == B.phi_blocks_elems
    assert np.array_equal(B.unpack(xb), np.asarray(phi, dtype=xb.dtype) * mask)
    yb = B.apply_blocks(xb)
    assert np.abs(B.unpack(yb) - yo).max() <= 1e-12 * np.abs(yo).max()
    # bit-identical to the dense-boundary call (same device path, different (un)packing)
    assert np.array_equal(yb, B.pack(B.apply(phi)))
    # unsharded context: the slice form carries the whole vector and is the same call
    assert B.phi_blocks_slice() == (0, xb.size)
    assert np.array_equal(B.apply_blocks_sharded(xb), yb)
    B.close()


def check_mps_boundary_paths(ctx, case, monkeypatch):
    """The MPS crosses the boundary as dense arrays; large site tensors are (un)packed by a strided-copy launch on the device,
    small ones by a host loop (CB200_PACK_MIN).  Both paths, sorted and interleaved link labels: bit-identical results."""
    st, bc, n, cp, q, nst, product = case
    site, W, qk = build_model(st, bc, n, cp)
    psi0 = warm_start(site, W, n, q, seed=9, maxdim=12)
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    sw = cb.Sweeps(2, [16, 24], 1e-10)
    out = {}
    for tag, pack_min in (("device", "1"), ("host", str(10 ** 15))):
        monkeypatch.setenv("CB200_PACK_MIN", pack_min)
        en, psi = cb.dmrg(H, [], to_cb(psi0), sw, ctx=ctx)
        ee = cb.calc_ee(psi, site.charges, ctx=ctx)
        out[tag] = (en, psi, ee)
    assert out["device"][0] == out["host"][0]
    assert all(np.array_equal(a, b) for a, b in zip(out["device"][1].sites, out["host"][1].sites))
    assert all(np.array_equal(a, b) for a, b in zip(out["device"][1].link_charges, out["host"][1].link_charges))
    assert np.array_equal(out["device"][2], out["host"][2])
    # interleaved labels on the inner links (many short sectors): permute the link index of a copy of the start state
    rng = np.random.default_rng(2)
    m = to_cb(psi0)
    for j in range(1, n):
        p = rng.permutation(len(m.link_charges[j]))
        m.link_charges[j] = np.ascontiguousarray(m.link_charges[j][p])
        m.sites[j - 1] = np.asfortranarray(m.sites[j - 1][:, :, p])
        m.sites[j] = np.asfortranarray(m.sites[j][p])
    monkeypatch.setenv("CB200_PACK_MIN", "1")
    en_p, _ = cb.dmrg(H, [], m, sw, ctx=ctx)
    assert abs(en_p - out["host"][0]) <= 1e-12 * abs(en_p)
    monkeypatch.delenv("CB200_PACK_MIN")


def check_heig_batched(ctx, cplx, sizes):
    """Batched Hermitian eigensolver (charge blocks of one bond side by side) against numpy.eigh, matrix by matrix."""
    rng = np.random.default_rng(sum(n * 7 + m for n, m in sizes) + cplx)
    As = []
    for n, m in sizes:
        X = (rng.standard_normal((n + 3, n)) + (1j * rng.standard_normal((n + 3, n)) if cplx else 0)) * (10.0 ** (-0.05 * np.arange(n)))
        A = X.conj().T @ X
        As.append(A / np.trace(A).real)
    out = cb.k_heig_batched(As, [m for _, m in sizes], ctx=ctx)
    for A, (n, m), (w, V) in zip(As, sizes, out):
        ev = np.linalg.eigvalsh(A)[::-1]
        assert np.abs(w - ev).max() <= 1e-14 * np.sqrt(n)
        if m:
            assert np.abs(V.conj().T @ V - np.eye(m)).max() <= 1e-13
            assert np.abs(A @ V - V @ (V.conj().T @ A @ V)).max() <= 2e-14 * np.sqrt(n)
            assert np.abs(np.sort(np.linalg.eigvalsh(V.conj().T @ A @ V))[::-1] - ev[:m]).max() <= 1e-14 * np.sqrt(n)


HEIG_BATCHES = [
    [(70, 30), (64, 64), (129, 5)],                   # three blocks of different sizes (different panel counts)
    [(200, 80), (1, 1), (2, 1), (333, 100)],          # tiny matrices ride along on the single-matrix path
    [(40, 10)] * 8,                                   # the largest batch
    [(33, 3)] * 9,                                    # more matrices than one batch takes: sequential
    [(515, 0), (130, 130)],                           # eigenvalues only for one of them
]


DMRG_CASES = [
    # site, bc, n, couplings, charge, nstates, product-state start?
    ("golden", "p", 6, [-1.0], 0, 3, True),
    ("golden", "o", 7, [-1.0], 0, 1, True),
    ("haagerup", "p", 6, [0.0, -1.0, 0.0], 0, 1, True),
    ("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 1, 1, False),   # real, charged sector
    ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 0, 1, False),   # complex128 block-sparse (README example model)
]


def check_dmrg_trajectory(ctx, case, nsweep=10):
    """Same start, same schedule => energies 1e-10 relative, EE 1e-8 (BASELINE.json tolerances).
    The first (unpenalised) state is compared bond by bond.  Penalised states (Weight = 1000) amplify a 1e-15
    perturbation to ~1e-7 in their transient even inside the oracle itself (tests/test_oracle.py::test_chaos_note),
    so for them the contract is checked on the converged last sweep and on the result."""
    st, bc, n, cp, q, nst, product = case
    site, W, qk = build_model(st, bc, n, cp)
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    rng = np.random.default_rng(5)
    done_o, done_g = [], []
    maxd = [10, 20, 40, 80]
    cut = [1e-5, 1e-6, 1e-7, 1e-8]
    for s in range(nst):
        if product:
            psi0 = od.product_mps(site, n, q, rng, dtype=complex if np.iscomplexobj(W[1]) else float)
        else:
            psi0 = warm_start(site, W, n, q, seed=7 + s)
        so = []
        eo, po = od.dmrg(W, done_o, psi0, od.Sweeps(nsweep, maxd, cut), weight=1000.0, stats=so)
        info = cb.DmrgInfo()
        eg, pg = cb.dmrg(H, done_g, to_cb(psi0), cb.Sweeps(nsweep, maxd, cut), weight=1000.0, ctx=ctx, info=info)
        assert len(so) == len(info.stats)
        for a, b in zip(so, info.stats):
            if s == 0 or a["sweep"] == nsweep - 1:
                assert a["m"] == b["m"], (a, b)
                assert abs(a["energy"] - b["energy"]) <= 1e-10 * max(1.0, abs(a["energy"])), (a, b)
        assert abs(eo - eg) <= 1e-10 * max(1.0, abs(eo))
        pgo = to_oracle(pg, site)
        assert abs(abs(od.overlap(po, pgo)) - 1) <= 1e-9
        # EE is first-order in the state (energy is second-order): 1e-8 needs the states equal to ~1e-8 in norm.  In the
        # charged Z3 sectors the lowest level is two-fold degenerate, the sweep converges slowly and a
        # rounding-level difference is amplified to a few 1e-8 in EE after 8-10 sweeps -- in the oracle against itself
        # too (different BLAS threading) -- so that case is held to 1e-7; every other case to BASELINE's 1e-8.
        ee_tol = 1e-7 if (st == "haagerupq" and q != 0) else 1e-8
        assert np.abs(np.array(od.calc_ee(po)) - np.array(od.calc_ee(pgo))).max() <= ee_tol
        assert abs(od.overlap(pgo, pgo) - 1) <= 1e-11
        done_o.append(po)
        done_g.append(pg)
    return done_g


def check_excited_qn(ctx, nsweep=12):
    """Penalised state in a Z3 sector (haagerupq K-only, q=0, second level -3.4971 non-degenerate).  Its sweep converges
    slowly (1e-10 needs ~50 sweeps, in the oracle too) and is chaotic in the transient, so only a loose agreement is
    asserted here; the exact checks of the penalty machinery are check_bond_ops (one bond) and the golden 3-state run."""
    st, bc, n, cp, q = "haagerupq", "p", 6, [-1.0, 0.0, 0.0], 0
    site, W, qk = build_model(st, bc, n, cp)
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    done_o, done_g, out = [], [], []
    for s in range(2):
        psi0 = warm_start(site, W, n, q, seed=7 + s)
        sw = ([10, 20, 40, 80], [1e-5, 1e-6, 1e-7, 1e-8])
        eo, po = od.dmrg(W, done_o, psi0, od.Sweeps(nsweep, *sw), weight=1000.0)
        eg, pg = cb.dmrg(H, done_g, to_cb(psi0), cb.Sweeps(nsweep, *sw), weight=1000.0, ctx=ctx)
        assert abs(eo - eg) <= (1e-10 if s == 0 else 2e-5) * abs(eo)
        assert abs(eg - ed_level(st, bc, n, cp, q)[s]) <= (1e-8 if s == 0 else 2e-3)
        pgo = to_oracle(pg, site)
        for prev in done_g:
            assert abs(od.overlap(to_oracle(prev, site), pgo)) <= 1e-3       # Weight = 1000 keeps it orthogonal
        done_o.append(po)
        done_g.append(pg)
        out.append(eg)
    return out

def check_excited_qn_converged(ctx, nsweep=100, measured=None):
    """The same penalised Z3 state as check_excited_qn, compared where north_star's tolerances are meant to hold: at CONVERGENCE.
    The sweep map contracts geometrically (factor ~0.72 per sweep: 1e-7 after 30 sweeps, 1e-13 after 70), so after 100 sweeps the engine
    and the oracle sit on the same fixed point whatever happened in the chaotic transient: energies 1e-10 relative (measured 1e-15),
    entanglement entropies 1e-8 (measured 8e-10), both levels equal to exact diagonalisation."""
    st, bc, n, cp, q = "haagerupq", "p", 6, [-1.0, 0.0, 0.0], 0
    site, W, qk = build_model(st, bc, n, cp)
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    done_o, done_g, out = [], [], []
    for s in range(2):
        psi0 = warm_start(site, W, n, q, seed=7 + s)
        sw = ([10, 20, 40, 80], [1e-5, 1e-6, 1e-7, 1e-8])
        eo, po = od.dmrg(W, done_o, psi0, od.Sweeps(nsweep, *sw), weight=1000.0)
        eg, pg = cb.dmrg(H, done_g, to_cb(psi0), cb.Sweeps(nsweep, *sw), weight=1000.0, ctx=ctx)
        pgo = to_oracle(pg, site)
        gaps = {"energy_rel": abs(eo - eg) / abs(eo), "ed_abs": abs(eg - ed_level(st, bc, n, cp, q)[s]),
                "overlap": abs(abs(od.overlap(po, pgo)) - 1),
                "ee": float(np.abs(np.array(od.calc_ee(po)) - np.array(od.calc_ee(pgo))).max()),
                "prev_overlap": max([abs(od.overlap(to_oracle(prev, site), pgo)) for prev in done_g], default=0.0)}
        if measured is not None:
            measured.append(gaps)
        assert gaps["energy_rel"] <= 1e-10, (s, eo, eg)
        assert gaps["ed_abs"] <= 1e-9, (s, eg)
        assert gaps["overlap"] <= 1e-9, gaps
        assert gaps["ee"] <= 1e-8, gaps
        assert gaps["prev_overlap"] <= 1e-6, gaps
        done_o.append(po)
        done_g.append(pg)
        out.append(eg)
    return out



def check_truncation_survives_ns_failure(ctx, case, monkeypatch):
    """split_bond's way out when heig_vectors cannot orthonormalise its inverse-iteration basis (forced: CB200_NS_MAXIT=-1 reports
    non-convergence whatever the basis looks like): the block's kept columns come from the one-sided Jacobi SVD of the gathered matrix instead.
    Same kept dimension per sector, same spectrum and truncation error, an isometry, and the same two-site tensor A.B as the default path."""
    st, bc, n, cp, q, bond = case
    site, W, qk, c = capture_bond(st, bc, n, cp, q, bond)
    B = make_bond(ctx, site, qk, c)
    phi = c["phi"] / np.linalg.norm(c["phi"])
    for d in ("left", "right"):
        monkeypatch.delenv("CB200_NS_MAXIT", raising=False)
        A0, B0, qm0, spec0, terr0 = B.truncate(phi, d, 40, 1e-8)
        monkeypatch.setenv("CB200_NS_MAXIT", "-1")
        A1, B1, qm1, spec1, terr1 = B.truncate(phi, d, 40, 1e-8)
        assert np.array_equal(qm0, qm1) and terr0 == terr1 and np.array_equal(spec0, spec1)     # decided before the vectors are formed
        iso = A1 if d == "left" else B1
        m = len(spec1)
        g = np.tensordot(iso.conj(), iso, axes=([0, 1], [0, 1])) if d == "left" else np.tensordot(iso, iso.conj(), axes=([1, 2], [1, 2]))
        assert np.abs(g - np.eye(m)).max() <= 1e-12
        assert np.abs(np.tensordot(A1, B1, axes=([2], [0])) - np.tensordot(A0, B0, axes=([2], [0]))).max() <= 1e-9
    monkeypatch.delenv("CB200_NS_MAXIT", raising=False)
    B.close()


def check_ground_state_kat(ctx, st, bc, n, cp, q, tol, nrep=12, seed=1):
    """Warm-up + reps of src/dmrg.h:538-609 through the C ABI; compare with the ED known answer."""
    site, W, qk = build_model(st, bc, n, cp)
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    rng = np.random.default_rng(seed)
    psi0 = od.product_mps(site, n, q, rng, dtype=complex if np.iscomplexobj(W[1]) else float)
    c = float(np.float32(1e-8))
    sw = cb.Sweeps(5, [10, 20, 40, 80, 200], [max(1e-5, c), max(1e-6, c), max(1e-7, c), max(1e-8, c), max(1e-9, c)])
    en, psi = cb.dmrg(H, [], to_cb(psi0), sw, weight=1000.0, ctx=ctx)
    for _ in range(nrep):
        en, psi = cb.dmrg(H, [], psi, cb.Sweeps(1, 400, c), weight=1000.0, ctx=ctx)
    e_ed = ed_level(st, bc, n, cp, q if site.nq > 1 else None)[0]
    assert abs(en - e_ed) <= tol * abs(e_ed), (en, e_ed)
    return en, psi, site


def check_packed_mps_boundary(ctx, case, shuffle=False):
    """cb200_mps.layout = 1 (ITensor QDense storage for the MPS of the coarse call): same run, bit for bit, as with dense arrays;
    previous states and the start state may arrive in different layouts; CalcEE and innerC take either."""
    st, bc, n, cp, q, nst, product = case
    site, W, qk = build_model(st, bc, n, cp)
    sq = site.charges
    psi0 = to_cb(warm_start(site, W, n, q, seed=13, maxdim=10))
    if shuffle:
        rng = np.random.default_rng(3)
        for j in range(1, n):
            p = rng.permutation(len(psi0.link_charges[j]))
            psi0.link_charges[j] = np.ascontiguousarray(psi0.link_charges[j][p])
            psi0.sites[j - 1] = np.asfortranarray(psi0.sites[j - 1][:, :, p])
            psi0.sites[j] = np.asfortranarray(psi0.sites[j][p])
        psi0.center = 0
    H = cb.MPO(W, qk, sq, nq=site.nq)
    sw = cb.Sweeps(2, [12, 20], 1e-10)
    pk0 = cb.PackedMPS.from_dense(psi0, sq)
    assert all(np.array_equal(a, b) for a, b in zip(pk0.to_dense().sites, psi0.sites))          # the packing itself loses nothing
    nelem = sum(int(((np.asarray(psi0.link_charges[j])[:, None, None] + np.asarray(sq)[None, :, None]
                      - np.asarray(psi0.link_charges[j + 1])[None, None, :]) % site.nq == 0).sum()) for j in range(n))
    assert sum(b.size for b in pk0.blocks) == nelem
    e_d, out_d = cb.dmrg(H, [], psi0, sw, ctx=ctx)
    e_p, out_p = cb.dmrg(H, [], pk0, sw, ctx=ctx)
    assert isinstance(out_p, cb.PackedMPS) and e_p == e_d
    assert all(np.array_equal(a, b) for a, b in zip(out_p.to_dense().sites, out_d.sites))
    assert all(np.array_equal(a, b) for a, b in zip(out_p.link_charges, out_d.link_charges))
    assert np.array_equal(cb.calc_ee(out_p, ctx=ctx), cb.calc_ee(out_d, sq, ctx=ctx))
    assert cb.overlap(out_p, out_p, ctx=ctx) == cb.overlap(out_d, out_d, sq, ctx=ctx)
    # an excited state with the penalty term: previous state packed, start state dense -- and the other way round
    e1, x1 = cb.dmrg(H, [out_p], psi0, sw, weight=1000.0, ctx=ctx)
    e2, x2 = cb.dmrg(H, [out_d], pk0, sw, weight=1000.0, ctx=ctx)
    assert e1 == e2 and all(np.array_equal(a, b) for a, b in zip(x2.to_dense().sites, x1.sites))
    # a REAL start state under a complex Hamiltonian: the packed data is widened at the boundary like the dense data
    if np.iscomplexobj(W[1]) and np.abs(np.imag(W[1])).max() > 0:
        pr = to_cb(od.product_mps(site, n, q, np.random.default_rng(4), dtype=float))
        assert not pr.is_complex()
        e3, x3 = cb.dmrg(H, [], pr, sw, ctx=ctx)
        e4, x4 = cb.dmrg(H, [], cb.PackedMPS.from_dense(pr, sq), sw, ctx=ctx)
        assert e3 == e4 and all(np.array_equal(a, b) for a, b in zip(x4.to_dense().sites, x3.sites))


TWO_STAGE_SIZES = [(300, 120), (257, 257), (449, 60), (641, 200)]


def check_heig_two_stage(ctx, cplx, n, m, monkeypatch):
    """The two-stage reduction (dense -> band by QR panels + GEMM block updates, band -> tridiagonal by bulge chasing, back-transformation
    through the chasing reflectors and the compact-WY panels) forced on at a small size, against numpy.eigh."""
    monkeypatch.setenv("CB200_TWOSTAGE_MIN_N", "256")
    rng = np.random.default_rng(n * 3 + m + cplx)
    X = (rng.standard_normal((n + 5, n)) + (1j * rng.standard_normal((n + 5, n)) if cplx else 0)) * (10.0 ** (-0.04 * np.arange(n)))
    X[:, 7] = X[:, 3]                                                      # an exactly rank-deficient direction
    A = X.conj().T @ X
    A /= np.trace(A).real
    w, V, it = cb.k_heig(A, m, ctx=ctx)
    ev = np.linalg.eigvalsh(A)[::-1]
    assert np.abs(w - ev).max() <= 1e-14 * np.sqrt(n)
    assert np.abs(V.conj().T @ V - np.eye(m)).max() <= 1e-13
    assert np.abs(A @ V - V @ (V.conj().T @ A @ V)).max() <= 2e-14 * np.sqrt(n)
    assert np.abs(np.sort(np.linalg.eigvalsh(V.conj().T @ A @ V))[::-1] - ev[:m]).max() <= 1e-14 * np.sqrt(n)
    # batched form: two-stage and one-stage matrices in the same call (the bulges of all two-stage blocks are chased in one launch)
    out = cb.k_heig_batched([A, A[:200, :200].copy(), A + 0.01 * np.eye(n)], [m, 50, m], ctx=ctx)
    assert np.abs(out[0][0] - ev).max() <= 1e-14 * np.sqrt(n)
    assert np.abs(out[2][0] - (ev + 0.01)).max() <= 1e-13 * np.sqrt(n)
    for (wi, Vi), Ai in zip(out, [A, A[:200, :200], A + 0.01 * np.eye(n)]):
        assert np.abs(Vi.conj().T @ Vi - np.eye(Vi.shape[1])).max() <= 1e-13
        assert np.abs(Ai @ Vi - Vi @ (Vi.conj().T @ Ai @ Vi)).max() <= 2e-14 * np.sqrt(n)


def check_true_svd_branch(ctx, cplx, direction="left", D=64):
    """ITensor's svdBond switches to a true SVD when noise == 0 and cutoff < 1e-12.  A two-site wavefunction with a
    graded singular spectrum sigma = 1 ... 1e-12 (sigma^2 down to 1e-24): the SVD branch must resolve sigma^2 far below eps * sigma_max^2,
    the denmat branch (cutoff >= 1e-12) cannot and need not; truncate() acts on sigma^2 in both."""
    rng = np.random.default_rng(5 + cplx)
    d, k = 2, 3
    n = D * d

    def rnd(*shape):
        return rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
    U, _ = np.linalg.qr(rnd(n, n))
    V, _ = np.linalg.qr(rnd(n, n))
    sig = 10.0 ** (-12.0 * np.arange(n) / (n - 1))
    X = (U * sig) @ V.conj().T                                           # rows (l, s1), columns (s2, r)
    X /= np.linalg.norm(X)
    sig = sig / np.linalg.norm(sig)
    phi = np.asfortranarray(X.reshape(D, d, d, D, order="F"))
    z = np.zeros(D, np.int32)
    W = np.asfortranarray(rng.standard_normal((k, d, d, k)))
    E = np.asfortranarray(rnd(D, k, D))                                   # complex environments make the bond problem complex128
    bond = cb.Bond(E + 0, W, W, E + 0, ql=z, qr=z, qkl=np.zeros(k, np.int32), qkm=np.zeros(k, np.int32), qkr=np.zeros(k, np.int32),
                   site_charges=np.zeros(d, np.int32), nq=1, ctx=ctx)
    p = np.linalg.svd(X, compute_uv=False) ** 2                           # LAPACK on the same stored matrix (rounding moved the smallest ones)
    assert np.abs(p / sig ** 2 - 1)[: n // 2].max() < 1e-6
    # ---- SVD branch: cutoff 1e-22 keeps every state whose discarded tail stays below 1e-22
    A, B, qm, spec, terr = bond.truncate(phi, direction, 10 ** 6, 1e-22)
    tail = np.cumsum(p[::-1])[::-1]                                       # tail[i] = sum_{j >= i} p_j
    m_want = int(np.sum(tail >= 1e-22))                                   # ITensor truncate: drop while truncerr + P(n) < cutoff
    assert abs(len(spec) - m_want) <= 1, (len(spec), m_want)
    got = np.sort(spec)[::-1]
    mm = min(len(got), m_want)
    rel = np.abs(got[:mm] / p[:mm] - 1)
    assert rel.max() <= 1e-4 and rel[p[:mm] > 1e-14].max() <= 1e-9, rel.max()        # RELATIVE accuracy down to sigma^2 ~ 1e-22
    AB = np.tensordot(A, B, axes=([2], [0]))
    assert np.abs(AB / np.linalg.norm(AB) - phi).max() <= 1e-11
    iso = A.reshape(n, -1, order="F") if direction == "left" else B.reshape(len(spec), n, order="F").T
    assert np.abs(iso.conj().T @ iso - np.eye(len(spec))).max() <= 1e-12
    # ---- denmat branch on the same input: same large values, nothing resolved below ~eps
    A2, B2, _, spec2, terr2 = bond.truncate(phi, direction, 10 ** 6, 1e-12)
    got2 = np.sort(spec2)[::-1]
    big = p[:len(got2)] > 1e-10
    assert np.abs(got2[big] / p[:len(got2)][big] - 1).max() <= 1e-5
    assert len(spec2) < len(spec)
    bond.close()
