"""Oracle (test infrastructure): term list -> compressed MPO, restating what
`toMPO(AutoMPO)` produces at src/golden_site.cpp:124, src/haagerup_site.cpp:236 and
src/haagerup_q_site.cpp:534.

ITensor's AutoMPO (third-party, not vendored) compresses, at every bond, the coefficient
matrix between the distinct operator strings to the left and to the right of the bond with
an SVD; the resulting bond dimension is 2 + rank. The same construction
is restated here: for cut b the crossing terms give C_b[alpha, beta] (alpha = ops on sites
<= b, beta = ops on sites > b), factored per Z_N flux block as C_b = X_b Y_b with Y_b having
orthonormal rows. MPO link state 0 is "nothing applied yet", the last state is "term
complete"; the channels in between carry the accumulated operator flux as their charge.

W[j] has index order (a, s_out, s_in, a'), link charge rule
q(a') = q(a) + q(s_out) - q(s_in)  (mod N).
"""
import numpy as np


def op_flux(O, charges, nq):
    """Z_N flux q(out) - q(in) of a site operator; raises if it is not homogeneous."""
    if nq == 1:
        return 0
    rows, cols = np.nonzero(np.abs(O) > 0)
    fl = set(((charges[r] - charges[c]) % nq) for r, c in zip(rows, cols))
    if len(fl) > 1:
        raise ValueError("operator is not charge-homogeneous")
    return int(fl.pop()) if fl else 0


def build_mpo(terms, site, n, tol=1e-12):
    """Return (W, qlink): W[j] (k_{j}, d, d, k_{j+1}) for j=0..n-1, qlink[b] the charges of cut b (b=0..n)."""
    d, nq, charges = site.d, site.nq, site.charges
    ops = site.ops
    flux = {nm: op_flux(O, charges, nq) for nm, O in ops.items()}
    cplx = any(np.iscomplexobj(ops[o]) and np.abs(ops[o].imag).max() > 0 for _, t in terms for o, _ in t) \
        or any(isinstance(c, complex) and c.imag != 0 for c, _ in terms)
    dt = complex if cplx else float

    # normalise terms: sorted by site, merged if identical
    merged = {}
    for c, t in terms:
        key = tuple(sorted((s, o) for o, s in t))
        assert len({s for s, _ in key}) == len(key), "two operators on one site in a term"
        merged[key] = merged.get(key, 0) + c
    merged = {k: v for k, v in merged.items() if v != 0}

    # per cut: rank factorisation of the crossing-coefficient matrix, blocked by flux of the left part
    X = [None] * (n + 1)   # X[b][alpha] -> vector over channels of cut b
    Y = [None] * (n + 1)   # Y[b][beta]  -> vector over channels of cut b
    qch = [None] * (n + 1)  # charges of channels of cut b
    for b in range(1, n):
        lefts, rights, entries = {}, {}, []
        for key, c in merged.items():
            if key[0][0] <= b < key[-1][0]:
                al = tuple(x for x in key if x[0] <= b)
                be = tuple(x for x in key if x[0] > b)
                ia = lefts.setdefault(al, len(lefts))
                ib = rights.setdefault(be, len(rights))
                entries.append((ia, ib, c))
        C = np.zeros((len(lefts), len(rights)), dtype=dt)
        for ia, ib, c in entries:
            C[ia, ib] += c
        lf = np.array([sum(flux[o] for _, o in al) % nq for al in lefts], dtype=np.int64)
        rf = np.array([(-sum(flux[o] for _, o in be)) % nq for be in rights], dtype=np.int64)
        xs = {al: [] for al in lefts}
        ys = {be: [] for be in rights}
        qs = []
        lkeys, rkeys = list(lefts), list(rights)
        for q in range(nq):
            ri = np.nonzero(lf == q)[0]
            ci = np.nonzero(rf == q)[0]
            if len(ri) == 0 or len(ci) == 0:
                continue
            Cq = C[np.ix_(ri, ci)]
            u, s, vh = np.linalg.svd(Cq, full_matrices=False)
            r = int((s > tol * max(s[0], 1e-300)).sum()) if s.size else 0
            Yq = vh[:r]                      # r x ncols, orthonormal rows
            Xq = Cq @ Yq.conj().T            # nrows x r
            assert np.abs(Xq @ Yq - Cq).max() < 1e-10 * max(1.0, np.abs(Cq).max())
            qs += [q] * r
            for a_ in lkeys:
                xs[a_].append(np.zeros(r, dtype=dt))
            for b_ in rkeys:
                ys[b_].append(np.zeros(r, dtype=dt))
            for k_, i in enumerate(ri):
                xs[lkeys[i]][-1] = Xq[k_]
            for k_, i in enumerate(ci):
                ys[rkeys[i]][-1] = Yq[:, k_]
        # anything with flux mismatch has zero coefficient by charge conservation
        for (ia, ib, c) in entries:
            assert lf[ia] == rf[ib], "term does not conserve charge"
        X[b] = {k_: (np.concatenate(v) if v else np.zeros(0, dtype=dt)) for k_, v in xs.items()}
        Y[b] = {k_: (np.concatenate(v) if v else np.zeros(0, dtype=dt)) for k_, v in ys.items()}
        qch[b] = np.array(qs, dtype=np.int64)

    kdim = [1] + [2 + len(qch[b]) for b in range(1, n)] + [1]
    qlink = [np.zeros(1, dtype=np.int64)]
    for b in range(1, n):
        qlink.append(np.concatenate([[0], qch[b], [0]]).astype(np.int64))
    qlink.append(np.zeros(1, dtype=np.int64))

    eye = np.eye(d)
    W = []
    for j in range(1, n + 1):
        kl, kr = kdim[j - 1], kdim[j]
        Wj = np.zeros((kl, d, d, kr), dtype=dt)
        start_l = 0
        end_l = kl - 1
        start_r = 0
        end_r = kr - 1
        if j < n:
            Wj[start_l, :, :, start_r] += eye
        if j > 1:
            Wj[end_l, :, :, end_r] += eye
        # single-site terms
        for key, c in merged.items():
            if len(key) == 1 and key[0][0] == j:
                Wj[start_l, :, :, end_r] += c * ops[key[0][1]]
        # terms starting at j
        if j < n:
            for al, x in X[j].items():
                if len(al) == 1 and al[0][0] == j:
                    for i in range(len(x)):
                        if x[i] != 0:
                            Wj[start_l, :, :, 1 + i] += x[i] * ops[al[0][1]]
        # terms ending at j
        if j > 1:
            for be, y in Y[j - 1].items():
                if len(be) == 1 and be[0][0] == j:
                    for i in range(len(y)):
                        if y[i] != 0:
                            Wj[1 + i, :, :, end_r] += y[i] * ops[be[0][1]]
        # terms passing through j
        if 1 < j < n:
            for bep, yp in Y[j].items():
                # right parts at cut j-1 that continue into bep at cut j: (j, o) + bep, or bep itself (identity on j)
                cands = [(bep, None)]
                for nm in ops:
                    cands.append((((j, nm),) + bep, nm))
                for be, nm in cands:
                    y = Y[j - 1].get(be)
                    if y is None:
                        continue
                    M = np.outer(y, yp.conj())
                    O = eye if nm is None else ops[nm]
                    nzi, nzj = np.nonzero(M)
                    for i, jj in zip(nzi, nzj):
                        Wj[1 + i, :, :, 1 + jj] += M[i, jj] * O
        Wj[np.abs(Wj) < 1e-15] = 0
        W.append(Wj)
    return W, qlink


def mpo_to_dense(W):
    """Contract an MPO to a dense d^n x d^n matrix (small n only); site 1 slowest."""
    T = W[0][0]            # (s_out, s_in, a')
    T = T.reshape(T.shape[0], T.shape[1], T.shape[2])
    for Wj in W[1:]:
        # T[(S_out), (S_in), a] * Wj[a, s_out, s_in, a']
        T = np.tensordot(T, Wj, axes=([2], [0]))        # S_out, S_in, s_out, s_in, a'
        so, si = T.shape[0], T.shape[1]
        T = T.transpose(0, 2, 1, 3, 4).reshape(so * Wj.shape[1], si * Wj.shape[2], Wj.shape[3])
    return T[:, :, 0]
