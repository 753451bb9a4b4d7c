"""Regenerates tests/golden/reference_models.json FROM THE REFERENCE'S OWN CODE and (with --ed) tests/golden/ed_levels_reference.json.

The reference's model files -- /root/reference/src/{golden_site,haagerup_site,haagerup_q_site,f_data}.cpp -- are compiled unmodified
(oracle/Makefile -> oracle/_ref/ref_models) against tests/itensor_stub, a miniature stand-in for the absent ITensor, and run here:
every site operator (`Op.set(...)` tables, src/golden_site.cpp:16-44, src/haagerup_site.cpp:25-109, src/haagerup_q_site.cpp:41-353), the
F-symbols and RhoDefectCell (src/f_data.cpp:6-111,242-270), state() labels, the Z3 sectors of the haagerupq site index
(src/haagerup_q_site.cpp:11-26) and the `mpo += ...` term lists of Hamiltonian() (src/golden_site.cpp:57-126,
src/haagerup_site.cpp:111-238, src/haagerup_q_site.cpp:355-536) for all five boundary conditions.  These are outputs of the reference
itself; tests/test_reference_pin.py asserts that oracle/models.py AND chain_b200/models.py reproduce them operator by operator and
term by term, and that the exact-diagonalisation levels of tests/golden/ed_levels.json follow from them.

  python tests/golden/from_reference.py          # rebuild + rerun the reference code, rewrite reference_models.json
  python tests/golden/from_reference.py --ed     # also rediagonalise H assembled from the dumped terms/operators (minutes)
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
FIXTURE = os.path.join(HERE, "reference_models.json")
ED_FIXTURE = os.path.join(HERE, "ed_levels_reference.json")


def run_reference(ref="/root/reference"):
    """Build oracle/_ref/ref_models from the reference sources and return its stdout (a JSON document)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "REF=" + ref], check=True, capture_output=True)
    return subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_models")], check=True, capture_output=True, text=True).stdout


def load_fixture():
    with open(FIXTURE) as f:
        return json.load(f)


# ---- H assembled from NOTHING but the dumped data (independent of oracle/models.py and chain_b200/models.py)
def ref_operator(dump, site, name):
    """d x d matrix O[out, in]: the dump holds M[i][j] = Op.elt(s(i+1), s'(j+1)) = <j|O|i>."""
    import numpy as np
    m = np.array(dump[site]["ops"][name])
    return (m[..., 0] + 1j * m[..., 1]).T


def ref_sparse_h(dump, site, bc, n, couplings, U=2.0):
    import numpy as np
    import scipy.sparse as sp
    case = [h for h in dump[site]["hamiltonians"] if h["bc"] == bc and h["n"] == n and h["U"] == U and list(h["couplings"]) == list(couplings)]
    if not case:
        raise KeyError("no dumped Hamiltonian for %s %s n=%d %s" % (site, bc, n, couplings))
    d = dump[site]["d"]
    eye = sp.identity(d, format="csr")
    H = sp.csr_matrix((d ** n, d ** n), dtype=complex)
    cache = {}
    for t in case[0]["terms"]:
        key = tuple(sorted((s, o) for o, s in t["ops"]))
        if key not in cache:
            by_site = {s: sp.csr_matrix(ref_operator(dump, site, o)) for o, s in t["ops"]}
            assert len(by_site) == len(t["ops"])
            M = None
            for s in range(1, n + 1):
                f = by_site.get(s, eye)
                M = f if M is None else sp.kron(M, f, format="csr")
            cache[key] = M
        H = H + complex(*t["c"]) * cache[key]
    if abs(H.imag).max() < 1e-14 if H.nnz else True:
        H = H.real
    return H.tocsr()


def ref_levels(dump, site, bc, n, couplings, charge, k):
    import numpy as np
    import scipy.sparse.linalg as spla
    H = ref_sparse_h(dump, site, bc, n, couplings)
    if charge is not None and site == "haagerupq":
        q = np.zeros(1, dtype=np.int64)
        ch = np.array(dump[site]["sector_charges"])
        for _ in range(n):
            q = (q[:, None] + ch[None, :]).reshape(-1)
        idx = np.nonzero(q % dump[site]["qn_mod"] == charge)[0]
        H = H[idx][:, idx]
    if H.shape[0] <= 600:
        return np.linalg.eigvalsh(H.toarray())[:k]
    return np.sort(spla.eigsh(H, k=k, which="SA", tol=1e-13, ncv=max(40, 4 * k), return_eigenvectors=False))


ED_CASES = [   # the subset of make_ed_golden.py's cases whose Hamiltonians are in the dump
    ("golden", "p", 6, [-1.0], None, 5), ("golden", "p", 8, [-1.0], None, 5), ("golden", "o", 6, [-1.0], None, 4),
    ("haagerup", "p", 6, [0.0, -1.0, 0.0], None, 6),
    ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 0, 5), ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 1, 5), ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 2, 5),
    ("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 0, 5), ("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 1, 4), ("haagerupq", "p", 6, [0.0, 0.0, -1.0], 0, 4),
    ("haagerupq", "s", 6, [0.0, -1.0, 0.0], 0, 6),
]

if __name__ == "__main__":
    text = run_reference()
    json.loads(text)
    with open(FIXTURE, "w") as f:
        f.write(text)
    print("wrote", FIXTURE, len(text), "bytes")
    if "--ed" in sys.argv:
        dump = json.loads(text)
        out = {"U": 2.0, "source": "exact diagonalisation of H assembled from reference_models.json (terms and operators computed by the reference's own code)", "levels": []}
        for st, bc, n, cp, q, k in ED_CASES:
            w = ref_levels(dump, st, bc, n, cp, q, k)
            out["levels"].append(dict(site=st, bc=bc, n=n, couplings=cp, charge=q, levels=[float(x) for x in w]))
            print(st, bc, n, cp, q, w)
        with open(ED_FIXTURE, "w") as f:
            json.dump(out, f, indent=1)
