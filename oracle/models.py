"""Oracle (test infrastructure): site operators and Hamiltonian term lists.

Restates, as plain NumPy tables and Python term lists, what the reference's site
classes define:
  * F-symbols of the Haagerup-Izumi Z_nu categories   -- src/f_data.cpp:6-92,242-270
  * golden   site ops + Hamiltonian                    -- src/golden_site.cpp:16-44,57-126
  * haagerup site ops + Hamiltonian                    -- src/haagerup_site.cpp:25-109,111-238
  * haagerupq site ops + Hamiltonian (Z3-graded basis) -- src/haagerup_q_site.cpp:11-26,41-353,355-536

Operator matrices are stored as O[out, in]: the reference writes `Op.set(s_(i), sP(j), v)`
with index pair (dag(s), s'), i.e. the operator maps |i> to v|j>, so O[j-1, i-1] = v.
A term is (coef, [(opname, site), ...]) with 1-based sites exactly as the `mpo += ...`
lines list them (periodic wrap through mod(), src/golden_site.cpp:3-7).
"""
import math
import numpy as np


# --------------------------------------------------------------------------- F-symbols
class FData:
    """Haagerup-Izumi Z_nu fusion data (src/f_data.cpp:6-92). Objects are 1..2nu."""

    def __init__(self, nu):
        self.nu = nu
        self.rho = nu + 1
        self.rk = 2 * nu
        self.phi = (math.sqrt(4 + nu * nu) + nu) / 2
        self.phi_inv = 1 / self.phi
        self.sqrt_phi_inv = math.sqrt(self.phi_inv)

    def is_invertible(self, i):
        return i <= self.nu

    def dual(self, i):
        return 1 + ((self.nu + 1 - i) % self.nu) if i <= self.nu else i

    def fusion(self, a, b):
        nu, rho = self.nu, self.rho
        if a <= nu and b <= nu:
            return {1 + ((a + b - 2) % nu)}
        if a <= nu and b > nu:
            return {rho + ((a + b - 2) % nu)}
        if a > nu and b <= nu:
            return self.fusion(1 + ((rho - b) % nu), a)
        ans = {1 + ((nu + a - b) % nu)}
        ans.update(range(rho, self.rk + 1))
        return ans

    def has_fusion(self, i, j, k):
        return k in self.fusion(i, j)

    def qd(self, i):
        return 1.0 if i <= self.nu else self.phi

    def add(self, i, j):
        return self.rho + ((i + j - 1) % self.nu)

    def fsymbol_pattern(self, i, j, k, l, m, n):
        d = self.dual
        if not (self.has_fusion(i, j, m) and self.has_fusion(k, d(l), d(m))
                and self.has_fusion(d(l), i, d(n)) and self.has_fusion(j, k, n)):
            return 0.0
        inv = self.is_invertible
        if inv(i) or inv(j) or inv(k) or inv(l):
            return 1.0
        if inv(m) and inv(n):
            return self.phi_inv
        if inv(m) or inv(n):
            return self.sqrt_phi_inv
        rho = self.rho
        if i != rho:
            return self.fsymbol(rho, j, self.add(k, i - rho), l, m, n)
        if j != rho:
            return self.fsymbol(rho, rho, k, self.add(l, j - rho), m, n)
        if k != rho:
            return self.fsymbol(rho, rho, rho, self.add(l, rho - k), m, self.add(n, rho - k))
        if m != rho:
            return self.fsymbol(rho, rho, rho, l, rho, self.add(n, m - rho))
        raise RuntimeError("FSymbolPattern failed.")

    def fsymbol(self, i, j, k, l, m, n):
        return 0.0

    def rho_defect_cell(self):
        """T[i1, j1, i2, j2] of src/f_data.cpp:94-111 (0-based object labels)."""
        rk = self.rk
        T = np.zeros((rk, rk, rk, rk))
        for i1 in range(1, rk + 1):
            for j1 in range(1, rk + 1):
                for i2 in range(1, rk + 1):
                    for j2 in range(1, rk + 1):
                        f = self.fsymbol(self.rho, i1, self.rho, j2, j1, i2) * math.sqrt(self.qd(i1) / self.qd(i2))
                        T[i1 - 1, j1 - 1, i2 - 1, j2 - 1] = f
        return T


class GoldenFData(FData):
    def __init__(self):
        super().__init__(1)

    def fsymbol(self, i, j, k, l, m, n):  # src/f_data.cpp:245-250
        rho = self.rho
        if i == rho and j == rho and k == rho and m == rho and not self.is_invertible(l) and not self.is_invertible(n):
            return -self.phi_inv
        return self.fsymbol_pattern(i, j, k, l, m, n)


class HaagerupFData(FData):
    def __init__(self):
        super().__init__(3)
        s13 = math.sqrt(13)
        self.x = (2 - s13) / 3
        self.y1 = (5 - s13 - math.sqrt(6 * (1 + s13))) / 12
        self.y2 = (5 - s13 + math.sqrt(6 * (1 + s13))) / 12
        self.z = (1 + s13) / 6

    def fsymbol(self, i, j, k, l, m, n):  # src/f_data.cpp:259-270
        rho = self.rho
        if i == rho and j == rho and k == rho and m == rho and not self.is_invertible(l) and not self.is_invertible(n):
            return {8: self.x, 9: self.y1, 10: self.y2, 11: self.z, 12: self.y1}[l + n]
        return self.fsymbol_pattern(i, j, k, l, m, n)


# --------------------------------------------------------------------------- site types
def _proj(d, i):
    P = np.zeros((d, d))
    P[i - 1, i - 1] = 1.0
    return P


class GoldenSite:
    """src/golden_site.cpp:11-53. d=2, states 1 ("1") and 2 ("tau")."""
    name = "golden"
    d = 2
    charges = np.zeros(2, dtype=np.int64)   # no QN
    nq = 1

    def __init__(self):
        f = GoldenFData()
        pi, spi = f.phi_inv, f.sqrt_phi_inv
        self.ops = {
            "id": np.eye(2),
            "n1": _proj(2, 1),
            "nt": _proj(2, 2),
            "FF": np.array([[pi * pi, spi * pi], [spi * pi, pi]]),
        }

    @staticmethod
    def state(name):  # 0-based basis index; src/golden_site.cpp:46-53
        return 0 if name == "1" else 1


class HaagerupSite:
    """src/haagerup_site.cpp:10-109. d=6: 1,a,b,rho,a rho,b rho."""
    name = "haagerup"
    d = 6
    charges = np.zeros(6, dtype=np.int64)
    nq = 1
    _FF = {  # opname -> (projector, left, right), src/haagerup_site.cpp:59-107
        "FF1": (1, 4, 4), "FFa": (1, 5, 5), "FFb": (1, 6, 6),
        "Frr": (4, 4, 4), "Farar": (4, 5, 5), "Fbrbr": (4, 6, 6),
        "Frar": (4, 4, 5), "Farbr": (4, 5, 6), "Fbrr": (4, 6, 4),
        "Frbr": (4, 4, 6), "Farr": (4, 5, 4), "Fbrar": (4, 6, 5),
        "FFar_1": (5, 1, 4), "FFar_a": (5, 2, 5), "FFar_b": (5, 3, 6),
        "Far_rr": (5, 4, 4), "Far_arar": (5, 5, 5), "Far_brbr": (5, 6, 6),
        "Far_rar": (5, 4, 5), "Far_arbr": (5, 5, 6), "Far_brr": (5, 6, 4),
        "Far_rbr": (5, 4, 6), "Far_arr": (5, 5, 4), "Far_brar": (5, 6, 5),
    }

    def __init__(self):
        f = HaagerupFData()
        self.fdata = f
        self.ops = {"id": np.eye(6)}
        for i, nm in enumerate(["n1", "na", "nb", "nr", "nar", "nbr"], start=1):
            self.ops[nm] = _proj(6, i)
        for nm, (p, l, r) in self._FF.items():
            v = np.array([f.fsymbol(l, 4, 4, r, i, p) for i in range(1, 7)])
            self.ops[nm] = np.outer(v, v)   # src/haagerup_site.cpp:32-41

    @staticmethod
    def state(name):  # always rho, src/haagerup_site.cpp:16-23
        return 3


class HaagerupQSite:
    """src/haagerup_q_site.cpp:11-353. d=6, Z3 charges (0,1,2,0,1,2) for states 1..6."""
    name = "haagerupq"
    d = 6
    charges = np.array([0, 1, 2, 0, 1, 2], dtype=np.int64)
    nq = 3

    def __init__(self):
        ops = {}
        ops["id"] = np.eye(6) / 3.0   # sic: src/haagerup_q_site.cpp:41-48
        third = 1.0 / 3

        def setop(pairs, dtype=float):
            O = np.zeros((6, 6), dtype=dtype)
            for (i, j, v) in pairs:
                O[j - 1, i - 1] = v
            return O

        ops["m1"] = setop([(1, 1, third), (2, 2, third), (3, 3, third)])
        ops["m2"] = setop([(1, 3, third), (2, 1, third), (3, 2, third)])
        ops["m3"] = setop([(1, 2, third), (2, 3, third), (3, 1, third)])
        ops["m4"] = setop([(4, 4, third), (5, 5, third), (6, 6, third)])
        ops["m5"] = setop([(4, 6, third), (5, 4, third), (6, 5, third)])
        ops["m6"] = setop([(4, 5, third), (5, 6, third), (6, 4, third)])
        a = 0.030557695601338686774
        b = 0.90832691319598393968
        c = 0.16660245292295808691
        ops["q1"] = setop([(1, 1, a), (2, 2, a), (3, 3, a), (4, 4, b), (1, 4, c), (4, 1, c)])
        ops["q2"] = setop([(1, 3, a), (2, 1, a), (3, 2, a), (2, 4, c), (4, 3, c)])
        ops["q3"] = setop([(1, 2, a), (2, 3, a), (3, 1, a), (3, 4, c), (4, 2, c)])

        def q_all(c_3, c_5, c_6, c_7, i, j):  # src/haagerup_q_site.cpp:107-248
            c_1 = 0.033641737525777182951
            c_2 = -0.018511383658106454101
            c_4 = 0.23240812075600178448
            c_8 = -0.10092521257733154885
            c3b, c5b, c6b, c7b = np.conj(c_3), np.conj(c_5), np.conj(c_6), np.conj(c_7)
            A = [(1, 1, c_1), (1, 4, c_2), (2, 2, c_1), (2, 5, c_3), (3, 3, c_1), (3, 6, c3b)]
            B = [(1, 3, c_1), (1, 6, c3b), (2, 1, c_1), (2, 4, c_2), (3, 2, c_1), (3, 5, c_3)]
            C = [(1, 2, c_1), (1, 5, c_3), (2, 3, c_1), (2, 6, c3b), (3, 1, c_1), (3, 4, c_2)]
            if (i, j) == (1, 1):
                P = A + [(4, 1, c_2), (4, 4, c_4), (5, 2, c3b), (5, 5, third), (6, 3, c_3), (6, 6, third)]
            elif (i, j) in ((1, 2), (2, 1)):
                P = B + [(4, 3, c_2), (4, 6, c_5), (5, 1, c3b), (5, 4, c_5), (6, 2, c_3), (6, 5, c_6)]
            elif (i, j) in ((1, 3), (3, 1)):
                P = C + [(4, 2, c_2), (4, 5, c5b), (5, 3, c3b), (5, 6, c6b), (6, 1, c_3), (6, 4, c5b)]
            elif (i, j) == (2, 2):
                P = C + [(4, 2, c_2), (4, 5, c_7), (5, 3, c3b), (5, 6, c5b), (6, 1, c_3), (6, 4, c_7)]
            elif (i, j) in ((2, 3), (3, 2)):
                P = A + [(4, 1, c_2), (4, 4, c_8), (5, 2, c3b), (6, 3, c_3)]
            elif (i, j) == (3, 3):
                P = B + [(4, 3, c_2), (4, 6, c7b), (5, 1, c3b), (5, 4, c7b), (6, 2, c_3), (6, 5, c_5)]
            else:
                raise KeyError((i, j))
            return setop(P, dtype=complex)

        qr = (-0.039825165312405310998 + 0.046388867673581487842j,
              -0.050462606288665774427 + 0.109830493882197205744j,
              -0.21966098776439441149j,
              0.16666666666666666667 + 0.14308449170979940122j)
        qar = (-0.020261355201913645545 - 0.057684038707248573062j,
               0.120347300956507040838 - 0.011213347953941672653j,
               -0.19023199562434830725 + 0.10983049388219720574j,
               -0.20724813804160352388 + 0.07279532144250674052j)
        qbr = (0.060086520514318956543 + 0.011295171033667085220j,
               -0.069884694667841266411 - 0.098617145928255533092j,
               0.19023199562434830725 + 0.10983049388219720574j,
               0.04058147137493685721 - 0.21587981315230614174j)
        for nm, cs in (("qr", qr), ("qar", qar), ("qbr", qbr)):
            for i in (1, 2, 3):
                for j in (1, 2, 3):
                    ops["%s%d%d" % (nm, i, j)] = q_all(*cs, i, j)
        self.ops = ops

    @staticmethod
    def state(name):  # src/haagerup_q_site.cpp:30-39
        return {"1": 4, "2": 5}.get(name, 3)


def make_site(site_type):
    return {"golden": GoldenSite, "haagerup": HaagerupSite, "haagerupq": HaagerupQSite}[site_type]()


# --------------------------------------------------------------------------- Hamiltonian term lists
def _wrap(x, n):
    return x - n if x > n else x


def _ranges(bc, n):
    """(L_U, L_K) of src/golden_site.cpp:68-74,100-106 (identical in the other site files)."""
    lu = n if bc == "p" else n - 1
    lk = n if bc == "p" else n - 2
    if bc == "sp":
        lu, lk = n - 2, n - 3
    return lu, lk


def golden_terms(bc, n, U, couplings):
    """src/golden_site.cpp:57-126."""
    K = couplings[0]
    lu, lk = _ranges(bc, n)
    terms = []
    if U != 0:
        for j in range(1, lu + 1):
            Uj = U * math.sin(math.pi * j / (lu + 1)) ** 2 if bc in ("s", "sp") else U
            terms.append((Uj, [("n1", j), ("n1", _wrap(j + 1, n))]))
        if bc == "d":
            UL = n * U
            terms += [(UL, [("id", 1)]), (UL, [("id", n)]), (-UL, [("nt", 1)]), (-UL, [("nt", n)])]
    if K != 0:
        for j in range(1, lk + 1):
            Kj = K * math.sin(math.pi * (j + 0.5) / (lk + 2)) ** 2 if bc in ("s", "sp") else K
            j1, j2 = _wrap(j + 1, n), _wrap(j + 2, n)
            terms.append((Kj, [("n1", j), ("nt", j1), ("n1", j2)]))
            terms.append((Kj, [("nt", j), ("FF", j1), ("nt", j2)]))
    return terms


def haagerup_terms(bc, n, U, couplings):
    """src/haagerup_site.cpp:111-238 (no sine-squared modulation for this site type)."""
    K, J, M = couplings[0], couplings[1], couplings[2]
    lu, lk = _ranges(bc, n)
    terms = []
    if U != 0:
        pairs = [("n1", "n1"), ("n1", "na"), ("n1", "nb"), ("n1", "nar"), ("n1", "nbr"),
                 ("na", "n1"), ("na", "na"), ("na", "nb"), ("na", "nr"), ("na", "nbr"),
                 ("nb", "n1"), ("nb", "na"), ("nb", "nb"), ("nb", "nr"), ("nb", "nar"),
                 ("nr", "na"), ("nr", "nb"),
                 ("nar", "n1"), ("nar", "nb"),
                 ("nbr", "n1"), ("nbr", "na")]
        for j in range(1, lu + 1):
            for a, b in pairs:
                terms.append((U, [(a, j), (b, _wrap(j + 1, n))]))
        if bc == "d":
            UL = n * U
            terms += [(UL, [("id", 1)]), (UL, [("id", n)]), (-UL, [("nr", 1)]), (-UL, [("nr", n)])]

    def triples(coef, lst):
        for j in range(1, lk + 1):
            j1, j2 = _wrap(j + 1, n), _wrap(j + 2, n)
            for a, b, c in lst:
                terms.append((coef, [(a, j), (b, j1), (c, j2)]))

    if K != 0:
        triples(K, [("n1", "nr", "n1"), ("na", "nar", "na"), ("nb", "nbr", "nb"),
                    ("nr", "FF1", "nr"), ("nar", "FFa", "nar"), ("nbr", "FFb", "nbr")])
    if J != 0:
        triples(J, [("n1", "nr", "nr"), ("nr", "nr", "n1"), ("na", "nar", "nar"), ("nar", "nar", "na"),
                    ("nb", "nbr", "nbr"), ("nbr", "nbr", "nb"),
                    ("nr", "Frr", "nr"), ("nr", "Frar", "nar"), ("nr", "Frbr", "nbr"),
                    ("nar", "Farr", "nr"), ("nar", "Farar", "nar"), ("nar", "Farbr", "nbr"),
                    ("nbr", "Fbrr", "nr"), ("nbr", "Fbrar", "nar"), ("nbr", "Fbrbr", "nbr")])
    if M != 0:
        triples(M, [("n1", "nr", "nar"), ("nar", "nr", "n1"), ("na", "nar", "nbr"), ("nbr", "nar", "na"),
                    ("nb", "nbr", "nr"), ("nr", "nbr", "nb"),
                    ("nr", "Far_rr", "nr"), ("nr", "Far_rar", "nar"), ("nr", "Far_rbr", "nbr"),
                    ("nar", "Far_arr", "nr"), ("nar", "Far_arar", "nar"), ("nar", "Far_arbr", "nbr"),
                    ("nbr", "Far_brr", "nr"), ("nbr", "Far_brar", "nar"), ("nbr", "Far_brbr", "nbr")])
    return terms


def haagerupq_terms(bc, n, U, couplings):
    """src/haagerup_q_site.cpp:355-536."""
    K, J, M = couplings[0], couplings[1], couplings[2]
    lu, lk = _ranges(bc, n)
    terms = []
    ssd = bc in ("s", "sp")
    if U != 0:
        for j in range(1, lu + 1):
            Uj = U * math.sin(math.pi * j / (lu + 1)) ** 2 if ssd else U
            j1 = _wrap(j + 1, n)
            for w, a, b in [(9, "m1", "m1"), (6, "m1", "m4"), (-3, "m2", "m6"), (-3, "m3", "m5"),
                            (6, "m4", "m1"), (-3, "m5", "m3"), (-3, "m6", "m2")]:
                terms.append((w * Uj, [(a, j), (b, j1)]))
        if bc == "d":
            UL = 3 * n * U
            terms += [(UL, [("id", 1)]), (UL, [("id", n)]), (-UL, [("m4", 1)]), (-UL, [("m4", n)])]

    def triples(base, lst):
        for j in range(1, lk + 1):
            cj = base * math.sin(math.pi * (j + 0.5) / (lk + 2)) ** 2 if ssd else base
            j1, j2 = _wrap(j + 1, n), _wrap(j + 2, n)
            for w, a, b, c in lst:
                terms.append((w * cj, [(a, j), (b, j1), (c, j2)]))

    if K != 0:
        triples(K, [(3, "m1", "m4", "m1"), (3, "m1", "m5", "m3"), (3, "m1", "m6", "m2"),
                    (3, "m2", "m4", "m3"), (3, "m2", "m5", "m2"), (3, "m2", "m6", "m1"),
                    (3, "m3", "m4", "m2"), (3, "m3", "m5", "m1"), (3, "m3", "m6", "m3"),
                    (3, "m4", "q1", "m4"), (3, "m4", "q2", "m6"), (3, "m4", "q3", "m5"),
                    (3, "m5", "q1", "m6"), (3, "m5", "q2", "m5"), (3, "m5", "q3", "m4"),
                    (3, "m6", "q1", "m5"), (3, "m6", "q2", "m4"), (3, "m6", "q3", "m6")])
    if J != 0:
        triples(J, [(3, "m1", "m4", "m4"), (3, "m1", "m5", "m6"), (3, "m1", "m6", "m5"),
                    (3, "m2", "m4", "m6"), (3, "m2", "m5", "m5"), (3, "m2", "m6", "m4"),
                    (3, "m3", "m4", "m5"), (3, "m3", "m5", "m4"), (3, "m3", "m6", "m6"),
                    (3, "m4", "m4", "m1"), (3, "m4", "m5", "m3"), (3, "m4", "m6", "m2"),
                    (3, "m5", "m4", "m3"), (3, "m5", "m5", "m2"), (3, "m5", "m6", "m1"),
                    (3, "m6", "m4", "m2"), (3, "m6", "m5", "m1"), (3, "m6", "m6", "m3"),
                    (9, "m4", "qr11", "m4"), (9, "m4", "qr12", "m6"), (9, "m4", "qr13", "m5"),
                    (9, "m5", "qr31", "m4"), (9, "m5", "qr32", "m6"), (9, "m5", "qr33", "m5"),
                    (9, "m6", "qr21", "m4"), (9, "m6", "qr22", "m6"), (9, "m6", "qr23", "m5")])
    if M != 0:
        w = -0.5 + math.sqrt(3) / 2 * 1j
        wb = -0.5 - math.sqrt(3) / 2 * 1j
        triples(M, [(3, "m1", "m4", "m4"), (3 * wb, "m1", "m5", "m6"), (3 * w, "m1", "m6", "m5"),
                    (3 * wb, "m2", "m4", "m6"), (3 * w, "m2", "m5", "m5"), (3, "m2", "m6", "m4"),
                    (3 * w, "m3", "m4", "m5"), (3, "m3", "m5", "m4"), (3 * wb, "m3", "m6", "m6"),
                    (3, "m4", "m4", "m1"), (3, "m4", "m5", "m3"), (3, "m4", "m6", "m2"),
                    (3 * w, "m5", "m4", "m3"), (3 * w, "m5", "m5", "m2"), (3 * w, "m5", "m6", "m1"),
                    (3 * wb, "m6", "m4", "m2"), (3 * wb, "m6", "m5", "m1"), (3 * wb, "m6", "m6", "m3"),
                    (9, "m4", "qar11", "m4"), (9, "m4", "qar12", "m6"), (9, "m4", "qar13", "m5"),
                    (9, "m5", "qar31", "m4"), (9, "m5", "qar32", "m6"), (9, "m5", "qar33", "m5"),
                    (9, "m6", "qar21", "m4"), (9, "m6", "qar22", "m6"), (9, "m6", "qar23", "m5")])
    return terms


def hamiltonian_terms(site_type, bc, n, U, couplings):
    f = {"golden": golden_terms, "haagerup": haagerup_terms, "haagerupq": haagerupq_terms}[site_type]
    return f(bc, n, U, list(couplings))


def z3_fourier_matrix():
    """haagerupq -> haagerup basis change Z[out(haagerup), in(haagerupq)], src/chain.cpp:311-335."""
    nrm = 1 / math.sqrt(3)
    w = -0.5 + math.sqrt(3) / 2 * 1j
    wb = -0.5 - math.sqrt(3) / 2 * 1j
    F3 = nrm * np.array([[1, 1, 1], [1, w, wb], [1, wb, w]])
    Z = np.zeros((6, 6), dtype=complex)
    Z[:3, :3] = F3
    Z[3:, 3:] = F3
    return Z
