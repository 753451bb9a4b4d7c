"""Oracle (test infrastructure): dense NumPy restatement of the two-site DMRG engine
the reference calls at src/dmrg.h:544,563,596 (`itensor::dmrg(H, psis, psi0, sweeps, args)`).

The engine itself is ITensor C++ v3 (third-party, un-vendored, un-pinned; absent from
/root/reference). Its published algorithm is restated here following SURVEY.md App. A:
  A.1 DMRGWorker sweep order       A.2 LocalMPO position/product
  A.3 LocalMPO_MPS (Weight penalty) A.4 davidson (MaxIter=niter)
  A.5 svdBond -> denmatDecomp / svd, truncate()
and the driver schedule of src/dmrg.h:527-653, CalcEE of src/utility.cpp:5-27.
PARITY UNPINNED against ITensor itself (no golden vectors exist); pinned by ED known answers.

Conventions: MPS tensor A[l, s, r]; MPO tensor W[a, s_out, s_in, a']; environments
E[x', a, x] (bra link, MPO link, ket link). Z_N symmetric runs are done with dense arrays
carrying a charge label per link value: dense arithmetic keeps forbidden blocks exactly zero,
and only the truncation (per-charge-block eigendecomposition, pooled selection) looks at labels.
"""
import numpy as np


class MPS:
    def __init__(self, A, qlink, site_charges, nq, center=None):
        self.A = A                      # list of (Dl, d, Dr) arrays
        self.q = qlink                  # list of n+1 int arrays: charge of each link value
        self.site_charges = np.asarray(site_charges)
        self.nq = nq
        self.center = center            # 1-based orthogonality centre, None if not canonical

    @property
    def n(self):
        return len(self.A)

    def copy(self):
        return MPS([a.copy() for a in self.A], [q.copy() for q in self.q], self.site_charges, self.nq, self.center)

    def bond_dims(self):
        return [a.shape[2] for a in self.A[:-1]]


def product_mps(site, n, charge, rng, dtype=float):
    """randomMPS(InitState) of src/dmrg.h:440-447 restated.
    Non-QN sites: random d-vector per site. QN sites: the single allowed element per site
    (state "0" everywhere, centre site = charge) with a random amplitude."""
    d = site.d
    A, q = [], [np.zeros(1, dtype=np.int64)]
    center = (n + 1) // 2
    for j in range(1, n + 1):
        v = np.zeros(d, dtype=dtype)
        if site.nq == 1:
            v[:] = rng.standard_normal(d)
            qs = 0
        else:
            st = site.state(str(charge)) if j == center else site.state("0")
            v[st] = rng.standard_normal()
            qs = int(site.charges[st])
        A.append(v.reshape(1, d, 1).copy())
        q.append((q[-1] + qs) % site.nq)
    return MPS(A, q, site.charges, site.nq)


# ----------------------------------------------------------------------------- environments
def update_left(L, A, W):
    """L'[m', a', m] = sum conj(A[l',t,m']) W[a,t,s,a'] L[l',a,l] A[l,s,m]   (A.2 makeL)."""
    X = np.tensordot(L, A, axes=([2], [0]))            # l' a s m
    X = np.tensordot(X, W, axes=([1, 2], [0, 2]))      # l' m t a'
    return np.tensordot(A.conj(), X, axes=([0, 1], [0, 2])).transpose(0, 2, 1)   # (m' m a') -> (m' a' m)


def update_right(R, B, W):
    """R'[m', a, m] = sum conj(B[m',t,r']) W[a,t,s,a''] R[r',a'',r] B[m,s,r]   (A.2 makeR)."""
    X = np.tensordot(R, B, axes=([2], [2]))            # r' a'' m s
    X = np.tensordot(X, W, axes=([1, 3], [3, 2]))      # r' m a t
    return np.tensordot(B.conj(), X, axes=([1, 2], [3, 0])).transpose(0, 2, 1)   # (m' m a) -> (m' a m)


def ovl_left(F, A, P):
    """F'[m, n] = sum conj(A[l,s,m]) F[l,j] P[j,s,n]  (overlap environment, A.2 Psi-flavour, conjugated)."""
    X = np.tensordot(F, P, axes=([1], [0]))            # l s n
    return np.tensordot(A.conj(), X, axes=([0, 1], [0, 1]))


def ovl_right(F, B, P):
    """F'[m, n] = sum conj(B[m,s,r]) F[r,p] P[n,s,p]."""
    X = np.tensordot(P, F, axes=([2], [1]))            # n s r
    return np.tensordot(B.conj(), X, axes=([1, 2], [1, 2]))


def heff_apply(L, W1, W2, R, phi):
    """(((phi*L)*W_b)*W_{b+1})*R  (A.2 product). phi[l,s1,s2,r] -> out[l',t1,t2,r']."""
    X = np.tensordot(L, phi, axes=([2], [0]))          # l' a s1 s2 r
    X = np.tensordot(X, W1, axes=([1, 2], [0, 2]))     # l' s2 r t1 b
    X = np.tensordot(X, W2, axes=([4, 1], [0, 2]))     # l' r t1 t2 c
    X = np.tensordot(X, R, axes=([4, 1], [1, 2]))      # l' t1 t2 r'
    return X


def heff_flops(Dl, Dr, d, kl, km, kr, cplx=False):
    """Dense flop count of the ITensor contraction order."""
    f = 2.0 * kl * Dl * Dl * d * d * Dr + 2.0 * kl * km * d ** 3 * Dl * Dr \
        + 2.0 * km * kr * d ** 3 * Dl * Dr + 2.0 * kr * Dr * Dr * d * d * Dl
    return f * (4 if cplx else 1)


# ----------------------------------------------------------------------------- Davidson (A.4)
RANDOMIZED = [0]   # number of times davidson() replaced a dependent residual by a random vector


def davidson(matvec, phi, maxiter=2, errgoal=1e-14, miniter=1, rng=None):
    size = phi.size
    actual_maxiter = min(maxiter, size - 1)
    nrm = np.linalg.norm(phi)
    if nrm == 0:
        phi = (rng or np.random.default_rng(0)).standard_normal(phi.shape).astype(phi.dtype)
        nrm = np.linalg.norm(phi)
    phi = phi / nrm
    V = [phi]
    AV = [matvec(phi)]
    lam = float(np.real(np.vdot(V[0], AV[0])))
    M = np.zeros((actual_maxiter + 2, actual_maxiter + 2), dtype=complex)
    M[0, 0] = lam
    last_lam = 1000.0
    nmv = 1
    for ii in range(actual_maxiter + 1):
        if ii == 0:
            q = AV[0] - lam * V[0]
        else:
            w, U = np.linalg.eigh(M[:ii + 1, :ii + 1])
            lam = float(w[0])
            u = U[:, 0]
            if np.real(u[0]) < 0:
                u = -u
            if not np.iscomplexobj(phi):
                u = np.real(u)
            phi = sum(u[k] * V[k] for k in range(ii + 1))
            q = sum(u[k] * AV[k] for k in range(ii + 1)) - lam * phi
        qnorm = np.linalg.norm(q)
        converged = (qnorm < errgoal and abs(lam - last_lam) < errgoal) or qnorm < max(1e-12, errgoal * 1e-3)
        if qnorm < 1e-20 or (converged and ii >= miniter) or ii == actual_maxiter:
            break
        last_lam = lam
        # one classical Gram-Schmidt pass (dots first, then subtractions)
        vq = [np.vdot(V[k], q) for k in range(ii + 1)]
        for k in range(ii + 1):
            q = q - vq[k] * V[k]
        qn = np.linalg.norm(q)
        passes = 1
        while qn < 1e-10 and passes <= 3:
            RANDOMIZED[0] += 1           # (test diagnostics: from here on the result depends on the generator, in ITensor too)
            q = (rng or np.random.default_rng(ii)).standard_normal(q.shape).astype(q.dtype)
            vq = [np.vdot(V[k], q) for k in range(ii + 1)]
            for k in range(ii + 1):
                q = q - vq[k] * V[k]
            qn = np.linalg.norm(q)
            passes += 1
        q = q / qn
        V.append(q)
        AV.append(matvec(q))
        nmv += 1
        for k in range(ii + 2):
            M[k, ii + 1] = np.vdot(V[k], AV[ii + 1])
            M[ii + 1, k] = np.conj(M[k, ii + 1])
    return lam, phi, nmv


# ----------------------------------------------------------------------------- truncation (A.5)
def truncate_probs(P, maxdim, mindim, cutoff):
    """ITensor `truncate()` restated: P sorted descending. Returns (m, truncerr, docut)."""
    P = np.array(P, dtype=float)
    origm = len(P)
    n = origm - 1
    zn = n
    while zn >= 0 and P[zn] < 0:
        P[zn] = 0
        zn -= 1
    if origm == 1:
        return 1, 0.0, P[0] / 2.0
    truncerr = 0.0
    while n >= maxdim:
        truncerr += P[n]
        n -= 1
    scale = P.sum()
    if scale == 0:
        scale = 1.0
    while n >= mindim and truncerr + P[n] < cutoff * scale:
        truncerr += P[n]
        n -= 1
    truncerr = truncerr / scale
    if n < 0:
        n = 0
    m = n + 1
    docut = 0.0
    if m < origm:
        docut = (P[m - 1] + P[m]) / 2.0
        if abs(P[m - 1] - P[m]) < 1e-3 * P[m - 1]:
            docut += 1e-3 * P[m - 1]
    return m, truncerr, docut


def split_bond(phi, ql, qr, site_charges, nq, direction, maxdim, mindim, cutoff, use_svd=False, normalize=True, diag=None):
    """svdBond of A.5. phi[l,s1,s2,r]; returns (A_b, A_{b+1}, qmid, spectrum, truncerr).
    direction 'left'  (Fromleft):  A_b is the isometry, A_{b+1} carries the weight (normalised).
    direction 'right' (Fromright): A_{b+1} is the isometry, A_b carries the weight."""
    Dl, d, _, Dr = phi.shape
    sq = np.asarray(site_charges)
    if direction == "left":
        Mx = phi.reshape(Dl * d, d * Dr)
        rowq = ((ql[:, None] + sq[None, :]) % nq).reshape(-1)        # charge of the new link for row (l,s1)
    else:
        Mx = phi.reshape(Dl * d, d * Dr).T                           # rows = (s2, r)
        rowq = ((qr[None, :] - sq[:, None]) % nq).reshape(-1)        # new link charge = q(r) - q(s2)
    blocks = []
    for q in range(nq):
        idx = np.nonzero(rowq == q)[0]
        if len(idx) == 0:
            continue
        Xq = Mx[idx]
        if use_svd:
            u, s, _ = np.linalg.svd(Xq, full_matrices=False)
            w = s * s
            if len(w) < len(idx):   # pad to the row dimension like an eigendecomposition would
                w = np.concatenate([w, np.zeros(len(idx) - len(w))])
                u = np.concatenate([u, np.zeros((len(idx), len(idx) - u.shape[1]), dtype=u.dtype)], axis=1)
        else:
            rho = Xq @ Xq.conj().T
            w, u = np.linalg.eigh(rho)
            w, u = w[::-1], u[:, ::-1]
        blocks.append((q, idx, w, u))
    pooled = np.sort(np.concatenate([b[2] for b in blocks]))[::-1]
    m, truncerr, docut = truncate_probs(pooled, maxdim, mindim, cutoff)
    if diag is not None:   # (test diagnostics: how close the cut runs to a multiplet / to zero-weight states)
        diag["cut_gap"] = float((pooled[m - 1] - pooled[m]) / pooled[m - 1]) if m < len(pooled) and pooled[m - 1] > 0 else 1.0
        diag["last_kept_rel"] = float(pooled[m - 1] / pooled[0]) if pooled[0] > 0 else 0.0
    Us, qmid, kept = [], [], []
    for q, idx, w, u in blocks:
        # ITensor's diagHImpl / svdImpl: a tensor WITHOUT QNs keeps exactly the m columns truncate() decided on (`reduceCols(UU, m)`); a QN
        # tensor keeps, block by block, the eigenvalues above the pooled threshold docut -- which drops a (near-)degenerate multiplet the cut
        # would have split (docut is raised by 1e-3 P(m-1) then) and any eigenvalue that is not positive.
        keep = min(m, len(w)) if nq == 1 else int((w > docut).sum())
        if keep == 0:
            continue
        Uq = np.zeros((Mx.shape[0], keep), dtype=phi.dtype)
        Uq[idx] = u[:, :keep]
        Us.append(Uq)
        qmid += [q] * keep
        kept.append(w[:keep])
    if not Us:   # degenerate corner: keep the single largest state
        q, idx, w, u = max(blocks, key=lambda b: b[2][0])
        Uq = np.zeros((Mx.shape[0], 1), dtype=phi.dtype)
        Uq[idx] = u[:, :1]
        Us, qmid, kept = [Uq], [q], [w[:1]]
    U = np.concatenate(Us, axis=1)
    mm = U.shape[1]
    qmid = np.array(qmid, dtype=np.int64)
    oc = U.conj().T @ Mx                     # (m, cols)
    if normalize:
        oc = oc / np.linalg.norm(oc)         # DoNormalize (svdBond); a plain svd() as in Swap keeps S*V as it is
    if direction == "left":
        A1 = U.reshape(Dl, d, mm)
        A2 = oc.reshape(mm, d, Dr)
    else:
        A2 = U.T.reshape(mm, d, Dr)          # B[m,(s2,r)] = U[(s2,r),m] (no conj: phi ~ oc^T-part * B)
        # phi[(l,s1),(s2,r)] = sum_m A1[(l,s1),m] B[m,(s2,r)] with B rows orthonormal => A1 = phi B^H
        A1 = (phi.reshape(Dl * d, d * Dr) @ A2.reshape(mm, d * Dr).conj().T)
        if normalize:
            A1 = A1 / np.linalg.norm(A1)
        A1 = A1.reshape(Dl, d, mm)
    return A1, A2, qmid, np.concatenate(kept), truncerr


# ----------------------------------------------------------------------------- the engine (A.1)
def orthogonalize_to_first(psi):
    """psi.position(1): right-orthogonalise sites n..2 (exact QR, no truncation)."""
    n = psi.n
    for j in range(n - 1, 0, -1):
        A = psi.A[j]
        Dl, d, Dr = A.shape
        Q, Rm = np.linalg.qr(A.reshape(Dl, d * Dr).T)       # (d Dr, k) (k, Dl)
        k = Q.shape[1]
        if k < Dl:
            # rank-deficient link: keep dimension k; charges follow the rows that survive
            raise NotImplementedError("link shrink in position(1) not needed for the supported starts")
        psi.A[j] = Q.T.reshape(k, d, Dr)
        psi.A[j - 1] = np.tensordot(psi.A[j - 1], Rm.T, axes=([2], [0]))
    psi.center = 1
    return psi


class Sweeps:
    """itensor::Sweeps restated: per-sweep maxdim/mindim/cutoff/niter/noise; short lists repeat their last entry."""

    def __init__(self, nsweep, maxdim, cutoff, niter=2, mindim=1, noise=0.0):
        def ext(x):
            x = list(x) if isinstance(x, (list, tuple)) else [x]
            return [x[min(i, len(x) - 1)] for i in range(nsweep)]
        self.nsweep = nsweep
        self.maxdim, self.cutoff, self.niter, self.mindim, self.noise = ext(maxdim), ext(cutoff), ext(niter), ext(mindim), ext(noise)


def dmrg(W, prev_states, psi0, sweeps, weight=1.0, stats=None, matvec_hook=None):
    """itensor::dmrg(H, psis, psi0, sweeps, {"Weight", w}) restated. Returns (energy, psi)."""
    psi = psi0.copy()
    n = psi.n
    nq, sq = psi.nq, psi.site_charges
    dt = np.result_type(*( [a.dtype for a in psi.A] + [w.dtype for w in W] + [a.dtype for p in prev_states for a in p.A]))
    psi.A = [a.astype(dt) for a in psi.A]
    if psi.center != 1:
        orthogonalize_to_first(psi)
    # environments: LE[b] = left env of cut b (sites 1..b), RE[b] = right env of cut b (sites b+1..n)
    LE = [None] * (n + 1)
    RE = [None] * (n + 1)
    LE[0] = np.ones((1, 1, 1), dtype=dt)
    RE[n] = np.ones((1, 1, 1), dtype=dt)
    FL = [[None] * (n + 1) for _ in prev_states]
    FR = [[None] * (n + 1) for _ in prev_states]
    for p in range(len(prev_states)):
        FL[p][0] = np.ones((1, 1), dtype=dt)
        FR[p][n] = np.ones((1, 1), dtype=dt)
    for b in range(n - 1, 1, -1):            # right envs for the first bond (sites 3..n)
        RE[b] = update_right(RE[b + 1], psi.A[b], W[b])
        for p, P in enumerate(prev_states):
            FR[p][b] = ovl_right(FR[p][b + 1], psi.A[b], P.A[b])
    energy = None
    for sw in range(sweeps.nsweep):
        maxdim, mindim, cutoff, niter = sweeps.maxdim[sw], sweeps.mindim[sw], sweeps.cutoff[sw], sweeps.niter[sw]
        use_svd = (sweeps.noise[sw] == 0 and cutoff < 1e-12)
        order = [(b, "left") for b in range(1, n)] + [(b, "right") for b in range(n - 1, 0, -1)]
        for b, direction in order:           # bond between sites b and b+1 (1-based)
            i = b - 1
            L, R = LE[b - 1], RE[b + 1]
            phi = np.tensordot(psi.A[i], psi.A[i + 1], axes=([2], [0]))
            proj = []
            for p, P in enumerate(prev_states):
                o = np.tensordot(FL[p][b - 1], P.A[i], axes=([1], [0]))
                o = np.tensordot(o, P.A[i + 1], axes=([2], [0]))
                o = np.tensordot(o, FR[p][b + 1], axes=([3], [1]))
                proj.append(o.astype(dt))
            W1, W2 = W[i], W[i + 1]

            def mv(x):
                y = heff_apply(L, W1, W2, R, x)
                for o in proj:
                    y = y + weight * np.vdot(o, x) * o
                return y
            if matvec_hook is not None:
                matvec_hook(dict(bond=b, dir=direction, L=L, W1=W1, W2=W2, R=R, phi=phi, proj=proj, ql=psi.q[b - 1], qr=psi.q[b + 1],
                                 sweep=sw, maxdim=maxdim, mindim=mindim, cutoff=cutoff, niter=niter, weight=weight))
            energy, phi, nmv = davidson(mv, phi, maxiter=niter)
            dg = {}
            A1, A2, qmid, spec, terr = split_bond(phi, psi.q[b - 1], psi.q[b + 1], sq, nq, direction,
                                                  maxdim, mindim, cutoff, use_svd=use_svd, diag=dg)
            psi.A[i], psi.A[i + 1], psi.q[b] = A1, A2, qmid
            if stats is not None:
                stats.append(dict(sweep=sw, bond=b, dir=direction, energy=energy, m=len(qmid), truncerr=terr,
                                  nmv=nmv, dims=(phi.shape[0], phi.shape[3]), cut_gap=dg["cut_gap"], last_kept_rel=dg["last_kept_rel"]))
            if direction == "left":
                LE[b] = update_left(L, A1, W1)
                for p, P in enumerate(prev_states):
                    FL[p][b] = ovl_left(FL[p][b - 1], A1, P.A[i])
            else:
                RE[b] = update_right(R, A2, W2)
                for p, P in enumerate(prev_states):
                    FR[p][b] = ovl_right(FR[p][b + 1], A2, P.A[i + 1])
    psi.A[0] = psi.A[0] / np.linalg.norm(psi.A[0])   # psi.normalize(); centre is at site 1
    psi.center = 1
    return energy, psi


# ----------------------------------------------------------------------------- measurements / driver
def calc_ee(psi):
    """CalcEE of src/utility.cpp:5-27: per-cut SVD spectrum, p > 1e-12 rule, natural log."""
    psi = psi.copy()
    n = psi.n
    orthogonalize_to_first(psi)
    out = []
    for b in range(1, n):
        A = psi.A[b - 1]
        Dl, d, Dr = A.shape
        u, s, vh = np.linalg.svd(A.reshape(Dl * d, Dr), full_matrices=False)
        p = s * s
        out.append(float(-(p[p > 1e-12] * np.log(p[p > 1e-12])).sum()))
        psi.A[b - 1] = u.reshape(Dl, d, -1)
        psi.A[b] = np.tensordot(s[:, None] * vh, psi.A[b], axes=([1], [0]))
    return out


def overlap(a, b):
    """<a|b> (innerC)."""
    E = np.ones((1, 1), dtype=complex)
    for x, y in zip(a.A, b.A):
        E = np.tensordot(np.tensordot(E, x.conj(), axes=([0], [0])), y, axes=([0, 1], [0, 1]))
    return E[0, 0]


def swap_translate(psi, cutoff=1e-5):
    """Measure("translation") of src/dmrg.h:855-858: Swap(psi, j) for j = 1..L-1 (src/chain.cpp:486-507), each a two-site swap
    gate followed by svd(wf, inds(psi(b)), {"Cutoff", kCutoff}) with U -> site b and S*V -> site b+1 (no renormalisation)."""
    psi = psi.copy()
    orthogonalize_to_first(psi)
    n = psi.n
    for b in range(n - 1):
        A, B = psi.A[b], psi.A[b + 1]
        wf = np.tensordot(A, B, axes=([2], [0])).transpose(0, 2, 1, 3)      # wf[l, s1, s2, r] = phi[l, s2, s1, r]
        Dl, d, _, Dr = wf.shape
        ql, qr = psi.q[b], psi.q[b + 2]
        A1, B1, qm, spec, terr = split_bond(np.ascontiguousarray(wf), ql, qr, psi.site_charges, psi.nq, "left", 10 ** 9, 1, cutoff,
                                            use_svd=True, normalize=False)
        psi.A[b], psi.A[b + 1] = A1, B1
        psi.q[b + 1] = np.asarray(qm)
    psi.center = n
    return psi


def rho_element(bra, G, ket, periodic=True):
    """<bra| rho |ket> with <s'|rho|s> = prod_j G[s_j, s'_j, s_{j+1}, s'_{j+1}] (RhoOp / OpenRhoOp, src/chain.cpp:65-185, written as a
    product of RhoDefectCells): transfer contraction carrying the first pair (ring closure) and the current pair."""
    d = G.shape[0]
    A, B = [np.asarray(x, dtype=complex) for x in ket.A], [np.asarray(x, dtype=complex) for x in bra.A]
    n = len(A)
    # X[i1, j1, i, j, m', m]: first pair, current pair, bra link, ket link
    X = np.zeros((d, d, d, d) + (B[0].shape[2], A[0].shape[2]), dtype=complex)
    for i in range(d):
        for j in range(d):
            X[i, j, i, j] = np.outer(B[0][0, j, :].conj(), A[0][0, i, :])
    for s in range(1, n):
        # X'[i1,j1,i,j,m',m] = sum_{ip,jp,x',x} X[i1,j1,ip,jp,x',x] G[ip,jp,i,j] conj(B[x',j,m']) A[x,i,m]
        X = np.einsum("abpqxy,pqij,xjm,yin->abijmn", X, G, B[s].conj(), A[s], optimize=True)
    if periodic:
        return np.einsum("abijmn,ijab->", X, G)          # ring closure G[p_n, p_1]; the outer links have dimension 1
    return X.sum()                                        # open chain: no closure factor


def full_analysis(energies, T, R, n, bc="p"):
    """Analyze, src/dmrg.h:1118-1174: m = E/L + T + R (periodic) | E/L + R (Dirichlet) | E/L; eigen(m); rotate; momentum and rho rounded
    to 1e-6; rows {energy, momentum, rho} sorted by energy."""
    E = np.diag(np.asarray(energies, dtype=complex))
    m = E / n + (T + R if bc == "p" else (R if bc == "d" else 0))
    w, U = np.linalg.eig(m)
    Ui = np.linalg.inv(U)
    Er, Tr, Rr = Ui @ E @ U, Ui @ T @ U, Ui @ R @ U
    out = []
    for i in range(len(energies)):
        mom = np.round((np.log(Tr[i, i]) / (2j * np.pi) * n).real * 1e6) / 1e6
        if mom == -(n // 2):
            mom = n // 2
        rho = np.round(Rr[i, i].real * 1e6) / 1e6 + 0.0
        out.append((Er[i, i].real, float(mom), float(rho)))
    return sorted(out)


def momentum_analysis(energies, T, n):
    """AnalyzeWithoutRho, src/dmrg.h:1177-1218: m = E/L + T, eigen(m), rotate E and T, momentum_i = round(Re(log(T_ii)/(2 pi i) L) 1e5)/1e5
    (with -L/2 -> L/2), rows sorted by energy."""
    E = np.diag(np.asarray(energies, dtype=complex))
    m = E / n + T
    w, U = np.linalg.eig(m)
    Ui = np.linalg.inv(U)
    Er, Tr = Ui @ E @ U, Ui @ T @ U
    out = []
    for i in range(len(energies)):
        mom = np.round((np.log(Tr[i, i]) / (2j * np.pi) * n).real * 1e5) / 1e5
        if mom == -(n // 2):
            mom = n // 2
        out.append((Er[i, i].real, float(mom)))
    return sorted(out)


def expect_mpo(psi, W):
    E = np.ones((1, 1, 1), dtype=complex)
    for A, Wj in zip(psi.A, W):
        E = update_left(E, A.astype(complex), Wj)
    return E[0, 0, 0]


def simulate(W, site, n, maxdim, cutoff, precision, nstates, charge=0, seed=0, weight=1000.0,
             max_reps=200, log=None, psi0=None):
    """The warm-up + rep schedule of DMRG::SimulateRaw (src/dmrg.h:527-653), without file I/O.
    cutoff/precision are taken as given (the CLI parses them as float: pass np.float32(x) to mimic)."""
    rng = np.random.default_rng(seed)
    cutoff = float(cutoff)
    precision = float(precision)
    done, energies, ee_first = [], [], None
    for st in range(nstates):
        init = psi0[st] if psi0 is not None else product_mps(site, n, charge, rng,
                                                              dtype=complex if np.iscomplexobj(W[1]) else float)
        sw = Sweeps(5, [10, 20, 40, 80, 200],
                    [max(1e-5, cutoff), max(1e-6, cutoff), max(1e-7, cutoff), max(1e-8, cutoff), max(1e-9, cutoff)], niter=2)
        en, psi = dmrg(W, done, init, sw, weight=weight)
        bond_dim, min_en, cnt, nsw = 200, 1e6, 3, 5
        for _ in range(max_reps):
            if st == 0:
                ee_first = calc_ee(psi)
            if bond_dim <= maxdim - 200:
                bond_dim += 200
            en, psi = dmrg(W, done, psi, Sweeps(1, bond_dim, cutoff, niter=2), weight=weight)
            nsw += 1
            if abs(min_en - en) * n < precision:
                cnt -= 1
            else:
                cnt, min_en = 3, en
            if log:
                log(st, nsw, en, max(psi.bond_dims()))
            if cnt == 0:
                break
        done.append(psi)
        energies.append(en)
    return energies, done, ee_first
