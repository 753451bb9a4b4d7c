"""Oracle (test infrastructure): exact diagonalisation of the restated Hamiltonians.

Plays the role of the reference's only built-in cross-check, mode 7 `Test()`
(src/dmrg.h:1283-1300: contract the whole MPO to a dense matrix and read off the lowest
levels). Here H is assembled as a sparse matrix directly from the term lists of
oracle/models.py, so it is independent of the MPO builder and of the DMRG restatement.
Site 1 is the slowest (left-most Kronecker factor).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .models import make_site, hamiltonian_terms


def build_sparse_h(site_type, bc, n, U, couplings):
    site = make_site(site_type)
    d = site.d
    terms = hamiltonian_terms(site_type, bc, n, U, couplings)
    cplx = any(np.iscomplexobj(site.ops[o]) for _, ops in terms for o, _ in ops) or \
        any(isinstance(c, complex) for c, _ in terms)
    H = sp.csr_matrix((d ** n, d ** n), dtype=complex if cplx else float)
    eye = sp.identity(d, format="csr")
    cache = {}
    for coef, ops in terms:
        key = tuple(sorted((s, o) for o, s in ops))
        if key not in cache:
            by_site = {}
            for o, s in ops:
                # two operators on one site never occur in the reference's term lists
                assert s not in by_site
                by_site[s] = sp.csr_matrix(site.ops[o])
            M = None
            for s in range(1, n + 1):
                f = by_site.get(s, eye)
                M = f if M is None else sp.kron(M, f, format="csr")
            cache[key] = M
        H = H + coef * cache[key]
    return H, site


def total_charge(site, n):
    """Total Z_N charge of every product basis state (site 1 slowest)."""
    d = site.d
    q = np.zeros(1, dtype=np.int64)
    for _ in range(n):
        q = (q[:, None] + site.charges[None, :]).reshape(-1)
    return q % site.nq


def lowest_levels(site_type, bc, n, U, couplings, k=6, charge=None):
    H, site = build_sparse_h(site_type, bc, n, U, couplings)
    if charge is not None and site.nq > 1:
        idx = np.nonzero(total_charge(site, n) == charge)[0]
        H = H[idx][:, idx]
    dim = H.shape[0]
    if dim <= 2000:
        # dense: a Krylov method returns one copy of a degenerate level unless rounding happens to split it -- how many copies of the 12-fold
        # lowest level of the open haagerupq chain (6^4 = 1296 states) came back depended on the BLAS thread count
        w = np.linalg.eigvalsh(H.toarray())
        return w[:k]
    w = spla.eigsh(H, k=k, which="SA", tol=1e-13, ncv=max(40, 4 * k), return_eigenvectors=False)
    return np.sort(w)


def ground_state(site_type, bc, n, U, couplings, charge=None):
    """Lowest eigenpair in the full (or charge-restricted) space; vector returned in the full basis."""
    H, site = build_sparse_h(site_type, bc, n, U, couplings)
    full = H.shape[0]
    idx = None
    if charge is not None and site.nq > 1:
        idx = np.nonzero(total_charge(site, n) == charge)[0]
        H = H[idx][:, idx]
    w, v = spla.eigsh(H, k=1, which="SA", tol=1e-14, ncv=40)
    vec = v[:, 0]
    if idx is not None:
        out = np.zeros(full, dtype=vec.dtype)
        out[idx] = vec
        vec = out
    return w[0], vec, site


def entanglement_entropies(vec, d, n):
    """S_b for cuts b=1..n-1 with the p>1e-12 rule of src/utility.cpp:18-24."""
    out = []
    for b in range(1, n):
        s = np.linalg.svd(vec.reshape(d ** b, d ** (n - b)), compute_uv=False)
        p = s * s
        p = p[p > 1e-12]
        out.append(float(-(p * np.log(p)).sum()))
    return out
