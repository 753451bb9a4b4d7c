"""Host-side model definitions feeding the hot path: site types, F-symbols, Hamiltonian terms and the MPO.

Mirrors the reference's site classes (same operator names, same couplings, same boundary-condition handling):
  Golden    -- /root/reference/src/golden_site.{h,cpp}
  Haagerup  -- /root/reference/src/haagerup_site.{h,cpp}
  HaagerupQ -- /root/reference/src/haagerup_q_site.{h,cpp}   (Z3-graded basis, charges 0,1,2,0,1,2)
  F-symbols -- /root/reference/src/f_data.{h,cpp}
`SiteSet.hamiltonian(bc, n, U, couplings)` plays the role of `sites_.Hamiltonian(...)` (src/dmrg.h:515) and returns
a chain_b200.MPO whose link dimensions equal what ITensor's toMPO(AutoMPO) yields (2 + rank of the per-bond
coefficient matrix), with Z3 charge labels on the MPO links for haagerupq.  Pure host code, O(L) work; the tensors
it produces are what cb200_dmrg() consumes.
"""
import cmath
import math
from collections import OrderedDict

import numpy as np

from .api import MPO, MPS

SQ13 = math.sqrt(13.0)


# --------------------------------------------------------------------------------------------- F-symbols (f_data.cpp)
class HaagerupIzumi:
    """Z_nu Haagerup-Izumi fusion data: objects 1..nu invertible, nu+1..2nu non-invertible (FData)."""

    def __init__(self, nu, special):
        self.nu, self.rho, self.rank = nu, nu + 1, 2 * nu
        self.qdim = (math.sqrt(4 + nu * nu) + nu) / 2
        self.special = special          # l+n -> F^{rho rho rho}_{l}(rho, n) for non-invertible l, n

    def inv(self, i):
        return i <= self.nu

    def dual(self, i):
        return 1 + (self.nu + 1 - i) % self.nu if self.inv(i) else i

    def fuse(self, a, b):
        nu, rho = self.nu, self.rho
        if self.inv(a) and self.inv(b):
            return {1 + (a + b - 2) % nu}
        if self.inv(a):
            return {rho + (a + b - 2) % nu}
        if self.inv(b):
            return self.fuse(1 + (rho - b) % nu, a)
        return {1 + (nu + a - b) % nu} | set(range(rho, self.rank + 1))

    def shift(self, i, j):
        return self.rho + (i + j - 1) % self.nu

    def F(self, i, j, k, l, m, n):
        rho = self.rho
        if i == j == k == m == rho and not self.inv(l) and not self.inv(n):
            return self.special(l + n)
        ok = (m in self.fuse(i, j) and self.dual(m) in self.fuse(k, self.dual(l))
              and self.dual(n) in self.fuse(self.dual(l), i) and n in self.fuse(j, k))
        if not ok:
            return 0.0
        if self.inv(i) or self.inv(j) or self.inv(k) or self.inv(l):
            return 1.0
        if self.inv(m) and self.inv(n):
            return 1.0 / self.qdim
        if self.inv(m) or self.inv(n):
            return math.sqrt(1.0 / self.qdim)
        if i != rho:
            return self.F(rho, j, self.shift(k, i - rho), l, m, n)
        if j != rho:
            return self.F(rho, rho, k, self.shift(l, j - rho), m, n)
        if k != rho:
            return self.F(rho, rho, rho, self.shift(l, rho - k), m, self.shift(n, rho - k))
        if m != rho:
            return self.F(rho, rho, rho, l, rho, self.shift(n, m - rho))
        raise ValueError("F-symbol pattern failed")


def golden_fdata():
    f = HaagerupIzumi(1, None)
    f.special = lambda s: -1.0 / f.qdim
    return f


def haagerup_fdata():
    x = (2 - SQ13) / 3
    y1 = (5 - SQ13 - math.sqrt(6 * (1 + SQ13))) / 12
    y2 = (5 - SQ13 + math.sqrt(6 * (1 + SQ13))) / 12
    z = (1 + SQ13) / 6
    tab = {8: x, 9: y1, 10: y2, 11: z, 12: y1}
    return HaagerupIzumi(3, lambda s: tab[s])


# --------------------------------------------------------------------------------------------- rho defect (measurement)
def rho_defect_cell(site_type):
    """G[i1, j1, i2, j2] = F(rho, i1, rho, j2; j1, i2) sqrt(QD(i1)/QD(i2)) -- FData::RhoDefectCell, src/f_data.cpp:94-111
    (i = ket / unprimed value, j = bra / primed value of the two neighbouring sites; 0-based here).  golden uses GoldenFData,
    haagerup and haagerupq (after the Z3 Fourier map) HaagerupFData (src/chain.cpp:118-131)."""
    f = golden_fdata() if site_type == "golden" else haagerup_fdata()
    rk = f.rank
    qd = [1.0 if f.inv(i) else f.qdim for i in range(1, rk + 1)]
    G = np.zeros((rk, rk, rk, rk))
    for i1 in range(1, rk + 1):
        for j1 in range(1, rk + 1):
            for i2 in range(1, rk + 1):
                for j2 in range(1, rk + 1):
                    G[i1 - 1, j1 - 1, i2 - 1, j2 - 1] = f.F(f.rho, i1, f.rho, j2, j1, i2) * math.sqrt(qd[i1 - 1] / qd[i2 - 1])
    return G


def rho_defect_mpos(site_type, n, periodic=True):
    """The rho-defect operator of RhoOp / OpenRhoOp (src/chain.cpp:65-185) as a SUM of MPOs with bond dimension d^2.

    Every MPO site tensor of the reference is B[j-1] * A[j] tied to the physical legs by copy tensors, i.e. the operator is a
    product of cells: <s'|rho|s> = prod_j G[s_j, s'_j, s_{j+1}, s'_{j+1}] (site n+1 = site 1 on a ring, absent on an open
    chain).  With the pair p_j = (s_j, s'_j) as the MPO bond this is W_j[p_{j-1}; s', s; p_j] = G[p_{j-1}, p_j] delta(p_j, (s, s'));
    the ring closure G[p_n, p_1] needs p_1 at the last site, so the ring operator is returned as one MPO per value of p_1
    (at most d^2 of them; their matrix elements add).  Returns a list of site-tensor lists [k_l, d(out), d(in), k_r]."""
    G = rho_defect_cell(site_type)
    d = G.shape[0]
    k = d * d
    pair = lambda i, j: i + d * j               # (ket value, bra value)

    def bulk():
        W = np.zeros((k, d, d, k))
        for ip in range(d):
            for jp in range(d):
                for i in range(d):
                    for j in range(d):
                        g = G[ip, jp, i, j]
                        if g != 0.0:
                            W[pair(ip, jp), j, i, pair(i, j)] = g
        return W
    Wb = bulk()
    mpos = []
    firsts = [(i1, j1) for i1 in range(d) for j1 in range(d)] if periodic else [None]
    for p1 in firsts:
        W1 = np.zeros((1, d, d, k))
        for i in range(d):
            for j in range(d):
                if p1 is None or (i, j) == p1:
                    W1[0, j, i, pair(i, j)] = 1.0
        Wn = np.zeros((k, d, d, 1))
        for ip in range(d):
            for jp in range(d):
                for i in range(d):
                    for j in range(d):
                        g = G[ip, jp, i, j] * (G[i, j, p1[0], p1[1]] if p1 is not None else 1.0)
                        if g != 0.0:
                            Wn[pair(ip, jp), j, i, 0] = g
        if not np.any(Wn):
            continue
        mpos.append([W1] + [Wb] * (n - 2) + [Wn])
    return mpos


def z3_fourier_matrix():
    """Z3FourierMatrix, src/chain.cpp:311-335: U[s_old, s_new] taking the haagerupq site basis to the haagerup one."""
    U = np.zeros((6, 6), dtype=complex)
    nrm = 1 / math.sqrt(3)
    om, omb = complex(-0.5, math.sqrt(3) / 2), complex(-0.5, -math.sqrt(3) / 2)
    for i in range(3):
        U[i, 0] = U[0, i] = nrm
        U[i + 3, 3] = U[3, i + 3] = nrm
    U[1, 1] = U[2, 2] = U[4, 4] = U[5, 5] = nrm * om
    U[1, 2] = U[2, 1] = U[4, 5] = U[5, 4] = nrm * omb
    return U


def to_haagerup_basis(psi):
    """removeQNs + Z3FourierTransform (src/dmrg.h:781-784, src/chain.cpp:337-345): a relabelling of the input of the rho
    measurement (site-local 6x6 unitary on each site tensor, link charges dropped), done on the host like the reference's
    other format conversions."""
    from .api import MPS
    U = z3_fourier_matrix()
    return MPS([np.einsum("lsr,st->ltr", np.asarray(a, dtype=complex), U) for a in psi.sites], None, nq=1, center=psi.center)


# --------------------------------------------------------------------------------------------- site sets
def _unit(d, pairs):
    """Operator from (in_state, out_state, value) triples, 1-based states; stored as O[out, in]."""
    cplx = any(isinstance(v, complex) for _, _, v in pairs)
    O = np.zeros((d, d), dtype=complex if cplx else float)
    for i, j, v in pairs:
        O[j - 1, i - 1] = v
    return O


class SiteSet:
    name = ""
    d = 0
    nq = 1
    charges = None

    def __init__(self, n):
        self.n = n
        self.ops = self._operators()

    def op(self, name):
        if name not in self.ops:
            raise KeyError("Operator name " + name + " not recognized")
        return self.ops[name]

    # -- the reference's Hamiltonian(boundary_condition, num_sites, U, couplings)
    def hamiltonian(self, bc, n, U, couplings, tol=1e-12):
        return build_mpo(self, self.terms(bc, n, U, list(couplings)), n, tol)

    @staticmethod
    def _spans(bc, n):
        if bc == "p":
            return n, n
        if bc == "sp":
            return n - 2, n - 3
        return n - 1, n - 2

    def init_state(self, charge=0, rng=None, dtype=float):
        """randomMPS(InitState(sites,"0") with the centre site set to `charge`) -- src/dmrg.h:440-447."""
        rng = rng or np.random.default_rng()
        n, d = self.n, self.d
        centre = (n + 1) // 2
        sites, q = [], [np.zeros(1, np.int32)]
        for j in range(1, n + 1):
            v = np.zeros(d, dtype=dtype)
            if self.nq == 1:
                v[:] = rng.standard_normal(d)
                qs = 0
            else:
                st = self.state(str(charge) if j == centre else "0")
                v[st] = rng.standard_normal()
                qs = int(self.charges[st])
            sites.append(v.reshape(1, d, 1))
            q.append(((q[-1] + qs) % self.nq).astype(np.int32))
        return MPS(sites, q, nq=self.nq, center=0)


class Golden(SiteSet):
    name, d, nq = "golden", 2, 1
    charges = np.zeros(2, np.int32)

    def _operators(self):
        f = golden_fdata()
        p = 1.0 / f.qdim
        return {"id": np.eye(2), "n1": _unit(2, [(1, 1, 1.0)]), "nt": _unit(2, [(2, 2, 1.0)]),
                "FF": np.array([[p * p, p * math.sqrt(p)], [p * math.sqrt(p), p]])}

    def state(self, s):
        return 0 if s == "1" else 1

    def terms(self, bc, n, U, couplings):
        K = couplings[0]
        nu, nk = self._spans(bc, n)
        w = lambda x: x - n if x > n else x
        ssd = bc in ("s", "sp")
        out = []
        if U != 0:
            for j in range(1, nu + 1):
                Uj = U * math.sin(math.pi * j / (nu + 1)) ** 2 if ssd else U
                out.append((Uj, (("n1", j), ("n1", w(j + 1)))))
            if bc == "d":
                out += [(n * U, (("id", 1),)), (n * U, (("id", n),)), (-n * U, (("nt", 1),)), (-n * U, (("nt", n),))]
        if K != 0:
            for j in range(1, nk + 1):
                Kj = K * math.sin(math.pi * (j + 0.5) / (nk + 2)) ** 2 if ssd else K
                out.append((Kj, (("n1", j), ("nt", w(j + 1)), ("n1", w(j + 2)))))
                out.append((Kj, (("nt", j), ("FF", w(j + 1)), ("nt", w(j + 2)))))
        return out


class Haagerup(SiteSet):
    name, d, nq = "haagerup", 6, 1
    charges = np.zeros(6, np.int32)
    LABEL = {1: "1", 2: "a", 3: "b", 4: "r", 5: "ar", 6: "br"}

    def _operators(self):
        f = haagerup_fdata()
        ops = {"id": np.eye(6)}
        for i, nm in self.LABEL.items():
            ops["n" + nm] = _unit(6, [(i, i, 1.0)])

        def ff(p, l, r):
            v = np.array([f.F(l, 4, 4, r, i, p) for i in range(1, 7)])
            return np.outer(v, v)
        # projector onto 1 in rho x rho (names FF1/FFa/FFb) and onto rho / a rho with all left/right labels
        for nm, x in (("FF1", 4), ("FFa", 5), ("FFb", 6)):
            ops[nm] = ff(1, x, x)
        for nm, g, x in (("FFar_1", 1, 4), ("FFar_a", 2, 5), ("FFar_b", 3, 6)):   # defined by the reference, unused by its Hamiltonian
            ops[nm] = ff(5, g, x)
        for l in (4, 5, 6):
            for r in (4, 5, 6):
                ops["F" + self.LABEL[l] + self.LABEL[r]] = ff(4, l, r)
                ops["Far_" + self.LABEL[l] + self.LABEL[r]] = ff(5, l, r)
        return ops

    def state(self, s):
        return 3

    def terms(self, bc, n, U, couplings):
        K, J, M = couplings[:3]
        nu, nk = self._spans(bc, n)
        w = lambda x: x - n if x > n else x
        L = self.LABEL
        out = []
        if U != 0:
            f = haagerup_fdata()
            # the 21 forbidden neighbour pairs: (x, y) with y not in x * rho  [written out pair by pair at src/haagerup_site.cpp:136-161]
            forbidden = [(x, y) for x in range(1, 7) for y in range(1, 7) if y not in f.fuse(x, 4)]
            for j in range(1, nu + 1):
                for x, y in forbidden:
                    out.append((U, (("n" + L[x], j), ("n" + L[y], w(j + 1)))))
            if bc == "d":
                out += [(n * U, (("id", 1),)), (n * U, (("id", n),)), (-n * U, (("nr", 1),)), (-n * U, (("nr", n),))]

        def tri(c, lst):
            for j in range(1, nk + 1):
                for a, b, cc in lst:
                    out.append((c, ((a, j), (b, w(j + 1)), (cc, w(j + 2)))))
        inv_pairs = [(1, 4), (2, 5), (3, 6)]          # invertible g and g*rho
        if K != 0:
            tri(K, [("n" + L[g], "n" + L[gr], "n" + L[g]) for g, gr in inv_pairs] +
                   [("n" + L[gr], "FF" + ("1" if g == 1 else L[g]), "n" + L[gr]) for g, gr in inv_pairs])
        if J != 0:
            lst = []
            for g, gr in inv_pairs:
                lst += [("n" + L[g], "n" + L[gr], "n" + L[gr]), ("n" + L[gr], "n" + L[gr], "n" + L[g])]
            lst += [("n" + L[l], "F" + L[l] + L[r], "n" + L[r]) for l in (4, 5, 6) for r in (4, 5, 6)]
            tri(J, lst)
        if M != 0:
            lst = []
            for g, gr in inv_pairs:
                nxt = 4 + (gr - 4 + 1) % 3          # a * (g rho)
                lst += [("n" + L[g], "n" + L[gr], "n" + L[nxt]), ("n" + L[nxt], "n" + L[gr], "n" + L[g])]
            lst += [("n" + L[l], "Far_" + L[l] + L[r], "n" + L[r]) for l in (4, 5, 6) for r in (4, 5, 6)]
            tri(M, lst)
        return out


class HaagerupQ(SiteSet):
    name, d, nq = "haagerupq", 6, 3
    charges = np.array([0, 1, 2, 0, 1, 2], np.int32)

    def _operators(self):
        t = 1.0 / 3
        ops = {"id": np.eye(6) / 3.0}
        # m1..m3: Z3 shifts on the invertible block, m4..m6 on the rho block (src/haagerup_q_site.cpp:50-84)
        for blk, base in ((0, 1), (3, 4)):
            ops["m%d" % base] = _unit(6, [(blk + i, blk + i, t) for i in (1, 2, 3)])
            ops["m%d" % (base + 1)] = _unit(6, [(blk + i, blk + 1 + (i + 1) % 3, t) for i in (1, 2, 3)])
            ops["m%d" % (base + 2)] = _unit(6, [(blk + i, blk + 1 + i % 3, t) for i in (1, 2, 3)])
        a, b, c = 0.030557695601338686774, 0.90832691319598393968, 0.16660245292295808691
        ops["q1"] = _unit(6, [(1, 1, a), (2, 2, a), (3, 3, a), (4, 4, b), (1, 4, c), (4, 1, c)])
        ops["q2"] = _unit(6, [(1, 3, a), (2, 1, a), (3, 2, a), (2, 4, c), (4, 3, c)])
        ops["q3"] = _unit(6, [(1, 2, a), (2, 3, a), (3, 1, a), (3, 4, c), (4, 2, c)])
        consts = {
            "qr": (-0.039825165312405310998 + 0.046388867673581487842j, -0.050462606288665774427 + 0.109830493882197205744j,
                   -0.21966098776439441149j, 0.16666666666666666667 + 0.14308449170979940122j),
            "qar": (-0.020261355201913645545 - 0.057684038707248573062j, 0.120347300956507040838 - 0.011213347953941672653j,
                    -0.19023199562434830725 + 0.10983049388219720574j, -0.20724813804160352388 + 0.07279532144250674052j),
            "qbr": (0.060086520514318956543 + 0.011295171033667085220j, -0.069884694667841266411 - 0.098617145928255533092j,
                    0.19023199562434830725 + 0.10983049388219720574j, 0.04058147137493685721 - 0.21587981315230614174j),
        }
        for nm, cs in consts.items():
            for i in (1, 2, 3):
                for j in (1, 2, 3):
                    ops["%s%d%d" % (nm, i, j)] = self._q_all(*cs, i, j)
        return ops

    @staticmethod
    def _q_all(c3, c5, c6, c7, i, j):
        """src/haagerup_q_site.cpp:107-248: the operator depends on (i,j) only through the shift (i+j) mod 3 of its
        upper block and a (i,j)-specific rho-block."""
        c1, c2, c4, c8 = 0.033641737525777182951, -0.018511383658106454101, 0.23240812075600178448, -0.10092521257733154885
        cj = lambda z: z.conjugate()
        sh = (i + j - 2) % 3      # 0: diagonal pattern, 1: "2-shift", 2: "1-shift"
        dst = {0: (1, 2, 3), 1: (3, 1, 2), 2: (2, 3, 1)}[sh]          # target of states 1,2,3 within the invertible block
        rho_coef = (c2, c3, cj(c3))                                      # coefficient to rho-block target 4,5,6
        P = []
        for s in (1, 2, 3):
            P.append((s, dst[s - 1], c1))
            tgt = 3 + dst[s - 1]
            P.append((s, tgt, rho_coef[tgt - 4]))
        low = {4: c2, 5: cj(c3), 6: c3}                                   # rho-block source -> invertible target coefficient
        rr = {  # rho-block -> rho-block coefficients per (i,j) class, listed for sources 4,5,6 (None = absent)
            (1, 1): (c4, 1.0 / 3, 1.0 / 3), (1, 2): (c5, c5, c6), (2, 1): (c5, c5, c6), (1, 3): (cj(c5), cj(c6), cj(c5)),
            (3, 1): (cj(c5), cj(c6), cj(c5)), (2, 2): (c7, cj(c5), c7), (2, 3): (c8, None, None), (3, 2): (c8, None, None),
            (3, 3): (cj(c7), cj(c7), c5),
        }[(i, j)]
        for k, s in enumerate((4, 5, 6)):
            P.append((s, dst[s - 4], low[s]))
            if rr[k] is not None:
                P.append((s, 3 + dst[s - 4], rr[k]))
        return _unit(6, [(a, b, complex(v)) for a, b, v in P])

    def state(self, s):
        return {"1": 4, "2": 5}.get(s, 3)

    def terms(self, bc, n, U, couplings):
        K, J, M = couplings[:3]
        nu, nk = self._spans(bc, n)
        w = lambda x: x - n if x > n else x
        ssd = bc in ("s", "sp")
        out = []
        if U != 0:
            for j in range(1, nu + 1):
                Uj = U * math.sin(math.pi * j / (nu + 1)) ** 2 if ssd else U
                for c, a, b in ((9, "m1", "m1"), (6, "m1", "m4"), (-3, "m2", "m6"), (-3, "m3", "m5"), (6, "m4", "m1"),
                                (-3, "m5", "m3"), (-3, "m6", "m2")):
                    out.append((c * Uj, ((a, j), (b, w(j + 1)))))
            if bc == "d":
                UL = 3 * n * U
                out += [(UL, (("id", 1),)), (UL, (("id", n),)), (-UL, (("m4", 1),)), (-UL, (("m4", n),))]

        def tri(base, lst):
            for j in range(1, nk + 1):
                cj_ = base * math.sin(math.pi * (j + 0.5) / (nk + 2)) ** 2 if ssd else base
                for c, a, b, cc in lst:
                    out.append((c * cj_, ((a, j), (b, w(j + 1)), (cc, w(j + 2)))))
        # shift bookkeeping: m1,m4 shift 0; m2,m5 shift "2"; m3,m6 shift "1" (flux of the operator); a triple must carry total flux 0
        inv = {0: "m1", 1: "m2", 2: "m3"}       # index = position in the reference's enumeration order
        rho = {0: "m4", 1: "m5", 2: "m6"}
        # the middle/right pairing used throughout src/haagerup_q_site.cpp:425-530: third index k = (-(i+j)) mod 3 in this enumeration
        third = lambda i, j: (-(i + j)) % 3
        qname = lambda stem, i, j: "%s%d%d" % (stem, i, j)
        if K != 0:
            lst = [(3, inv[i], rho[j], inv[third(i, j)]) for i in range(3) for j in range(3)]
            lst += [(3, rho[i], "q%d" % (j + 1), rho[third(i, j)]) for i in range(3) for j in range(3)]
            tri(K, lst)
        if J != 0:
            lst = [(3, inv[i], rho[j], rho[third(i, j)]) for i in range(3) for j in range(3)]
            lst += [(3, rho[i], rho[j], inv[third(i, j)]) for i in range(3) for j in range(3)]
            rowmap = {0: 1, 1: 3, 2: 2}         # m4 -> qr1*, m5 -> qr3*, m6 -> qr2*
            lst += [(9, rho[i], qname("qr", rowmap[i], j + 1), rho[(-j) % 3]) for i in range(3) for j in range(3)]
            tri(J, lst)
        if M != 0:
            om = cmath.exp(2j * math.pi / 3)
            omb = om.conjugate()
            ph = {0: 1.0, 1: om, 2: omb}
            # first block: coefficient omega^(-(j + 2k))-type phases, written out at src/haagerup_q_site.cpp:502-510
            first = [(0, 0, 1.0), (0, 1, omb), (0, 2, om), (1, 0, omb), (1, 1, om), (1, 2, 1.0), (2, 0, om), (2, 1, 1.0), (2, 2, omb)]
            lst = [(3 * c, inv[i], rho[j], rho[third(i, j)]) for i, j, c in first]
            lst += [(3 * ph[i], rho[i], rho[j], inv[third(i, j)]) for i in range(3) for j in range(3)]
            rowmap = {0: 1, 1: 3, 2: 2}
            lst += [(9, rho[i], qname("qar", rowmap[i], j + 1), rho[(-j) % 3]) for i in range(3) for j in range(3)]
            tri(M, lst)
        return out


def make_siteset(site_type, n):
    return {"golden": Golden, "haagerup": Haagerup, "haagerupq": HaagerupQ}[site_type](n)


# --------------------------------------------------------------------------------------------- toMPO(AutoMPO)
def _flux(O, charges, nq):
    if nq == 1:
        return 0
    r, c = np.nonzero(np.abs(O) > 0)
    f = {int((charges[i] - charges[j]) % nq) for i, j in zip(r, c)}
    if len(f) > 1:
        raise ValueError("operator without a definite Z_N flux")
    return f.pop() if f else 0


def build_mpo(sites, terms, n, tol=1e-12):
    """Compressed MPO of sum_t c_t prod_k O_k(site_k): per-cut rank factorisation C_b = X_b Y_b of the coefficient matrix
    between left and right operator strings (what ITensor's toMPO(AutoMPO) does; k_b = 2 + rank C_b)."""
    d, nq, q = sites.d, sites.nq, sites.charges
    opflux = {nm: _flux(O, q, nq) for nm, O in sites.ops.items()}
    acc = OrderedDict()
    for c, ops in terms:
        key = tuple(sorted((s, nm) for nm, s in ops))
        acc[key] = acc.get(key, 0) + c
    acc = OrderedDict((k, v) for k, v in acc.items() if v != 0)
    cplx = any(isinstance(v, complex) and v.imag != 0 for v in acc.values()) or \
        any(np.iscomplexobj(sites.ops[nm]) and np.abs(sites.ops[nm].imag).max() > 0 for k in acc for _, nm in k)
    dt = complex if cplx else float

    cut = [None] * (n + 1)      # per cut: dict(left string -> x vector), dict(right string -> y vector), channel charges
    for b in range(1, n):
        L, R, ent = OrderedDict(), OrderedDict(), []
        for key, c in acc.items():
            if key[0][0] <= b < key[-1][0]:
                le = tuple(x for x in key if x[0] <= b)
                ri = tuple(x for x in key if x[0] > b)
                ent.append((L.setdefault(le, len(L)), R.setdefault(ri, len(R)), c))
        C = np.zeros((len(L), len(R)), dtype=dt)
        for i, j, c in ent:
            C[i, j] += c
        lf = np.array([sum(opflux[nm] for _, nm in s) % nq for s in L], dtype=int)
        rf = np.array([(-sum(opflux[nm] for _, nm in s)) % nq for s in R], dtype=int)
        Xs, Ys, qs = [], [], []
        for f in range(nq):
            ri_, ci_ = np.nonzero(lf == f)[0], np.nonzero(rf == f)[0]
            if len(ri_) == 0 or len(ci_) == 0:
                continue
            Cf = C[np.ix_(ri_, ci_)]
            u, s, vh = np.linalg.svd(Cf, full_matrices=False)
            r = int((s > tol * max(s[0], 1e-300)).sum())
            Y = np.zeros((r, len(R)), dtype=dt)
            Y[:, ci_] = vh[:r]
            X = np.zeros((len(L), r), dtype=dt)
            X[ri_] = Cf @ vh[:r].conj().T
            Xs.append(X); Ys.append(Y); qs += [f] * r
        X = np.concatenate(Xs, axis=1) if Xs else np.zeros((len(L), 0), dtype=dt)
        Y = np.concatenate(Ys, axis=0) if Ys else np.zeros((0, len(R)), dtype=dt)
        if np.abs(X @ Y - C).max(initial=0) > 1e-10 * max(1.0, np.abs(C).max(initial=0)):
            raise ValueError("Hamiltonian terms do not conserve the Z_N charge")
        cut[b] = (OrderedDict((s, X[i]) for s, i in L.items()), OrderedDict((s, Y[:, j]) for s, j in R.items()),
                  np.array(qs, dtype=np.int32))

    k = [1] + [2 + len(cut[b][2]) for b in range(1, n)] + [1]
    qlink = [np.zeros(1, np.int32)] + [np.concatenate([[0], cut[b][2], [0]]).astype(np.int32) for b in range(1, n)] + [np.zeros(1, np.int32)]
    I = np.eye(d)
    W = []
    for j in range(1, n + 1):
        Wj = np.zeros((k[j - 1], d, d, k[j]), dtype=dt)
        if j < n:
            Wj[0, :, :, 0] += I
        if j > 1:
            Wj[-1, :, :, -1] += I
        for key, c in acc.items():                                   # on-site terms
            if len(key) == 1 and key[0][0] == j:
                Wj[0, :, :, -1] += c * sites.ops[key[0][1]]
        if j < n:                                                    # strings that start here
            for s, x in cut[j][0].items():
                if len(s) == 1 and s[0][0] == j:
                    Wj[0, :, :, 1:1 + len(x)] += sites.ops[s[0][1]][:, :, None] * x[None, None, :]
        if j > 1:                                                    # strings that end here
            for s, y in cut[j - 1][1].items():
                if len(s) == 1 and s[0][0] == j:
                    Wj[1:1 + len(y), :, :, -1] += y[:, None, None] * sites.ops[s[0][1]][None, :, :]
        if 1 < j < n:                                                # strings passing through (identity or an operator on j)
            Yl, Yr = cut[j - 1][1], cut[j][1]
            for s, yr in Yr.items():
                for nm, O in [(None, I)] + list(sites.ops.items()):
                    sl = s if nm is None else ((j, nm),) + s
                    yl = Yl.get(sl)
                    if yl is None:
                        continue
                    Wj[1:1 + len(yl), :, :, 1:1 + len(yr)] += np.einsum("a,ts,b->atsb", yl, O, yr.conj())
        Wj[np.abs(Wj) < 1e-15] = 0
        W.append(Wj)
    return MPO(W, qlink, sites.charges, nq=nq)


def mpo_to_dense(mpo):
    T = mpo.sites[0][0]
    for Wj in mpo.sites[1:]:
        T = np.tensordot(T, Wj, axes=([2], [0]))
        a, b = T.shape[0], T.shape[1]
        T = T.transpose(0, 2, 1, 3, 4).reshape(a * Wj.shape[1], b * Wj.shape[2], Wj.shape[3])
    return T[:, :, 0]
