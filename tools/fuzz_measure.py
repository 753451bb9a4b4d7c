"""Differential fuzz of the measurement entry points against DENSE state-vector arithmetic on RANDOM matrix-product states:
`cb200_mps_entropies` (CalcEE, src/utility.cpp:5-27), `cb200_mps_overlap` (innerC, src/dmrg.h:1063), `cb200_mps_translate` (Swap chain,
src/chain.cpp:486-507) and `cb200_mps_mpo_element` (<a|O|b>, the rho-defect element of src/dmrg.h:1016-1021).

The committed tests feed these with DMRG ground states of the documented models (canonical form, charge-sorted links, moderate bond dimensions).
Here the states are arbitrary flux-conserving tensors: any site type, length 2-7, link dimensions 1-9 drawn per link, link charge labels in RANDOM
(unsorted, interleaved) order, ragged and empty sectors, real or complex, not normalised, not in any canonical form, both storage layouts
(dense arrays and ITensor's block-packed QDense order).  Tolerances: entropies 1e-10, overlaps / matrix elements 1e-12 relative to the norms,
translation (run with cutoff 1e-24, i.e. exact: a relative cutoff c discards singular values up to sqrt(c), 1e-7 at 1e-14): the translated state equals the cyclically shifted state vector to 1e-9.  The reference values
come from the full d^n state vector (NumPy; no MPS algebra at all), so rank-deficient links -- which the oracle's position(1) does not
support -- are covered too.

    python tools/fuzz_measure.py [NDRAWS=200] [SEED0=0]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_mps(rng, site, n, total, cplx, dmax=9):
    """Random flux-conserving MPS: A_j[l, s, r] != 0 only where q(l) + q(s) = q(r) (mod nq); q(link 0) = 0, q(link n) = total."""
    from oracle import dmrg as od
    nq, sq = site.nq, np.asarray(site.charges)
    q = [np.zeros(1, np.int64)]
    for j in range(1, n):
        dim = int(rng.integers(1, dmax + 1))
        q.append(rng.integers(0, nq, size=dim).astype(np.int64))            # unsorted labels, sectors may be empty
    q.append(np.full(1, total % nq, np.int64))
    A = []
    for j in range(n):
        shape = (len(q[j]), len(sq), len(q[j + 1]))
        t = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
        mask = (q[j][:, None, None] + sq[None, :, None] - q[j + 1][None, None, :]) % nq == 0
        A.append(t * mask * float(rng.uniform(0.3, 2.0)))
    return od.MPS(A, q, site.charges, nq, center=None)


SHIFT = {}


def run_one(ctx, seed):
    import chain_b200 as cb
    import common
    from oracle import models
    rng = np.random.default_rng(50000 + seed)
    st = ["golden", "haagerup", "haagerupq"][rng.integers(3)]
    site = models.make_site(st)
    n = int(rng.integers(2, 8 if st == "golden" else 6))
    cplx = bool(rng.random() < 0.5)
    total = int(rng.integers(site.nq))
    a, b = random_mps(rng, site, n, total, cplx), random_mps(rng, site, n, total, cplx and rng.random() < 0.8)
    def full(m):                        # the state vector as a (d,)*n tensor
        v = m.A[0][0]                   # (d, r)
        for j in range(1, m.n):
            v = np.tensordot(v, m.A[j], axes=([v.ndim - 1], [0]))
        return v[..., 0]

    va, vb = full(a), full(b)
    na, nb = np.linalg.norm(va), np.linalg.norm(vb)
    if na == 0 or nb == 0:
        return "edge", "a random state vanished identically"
    packed = bool(rng.random() < 0.5)
    to = (lambda m: cb.PackedMPS.from_dense(common.to_cb(m), site.charges)) if packed else common.to_cb
    tag = "%s n=%d cplx=%s q=%d packed=%s dims=%s" % (st, n, cplx, total, packed, [len(x) for x in a.q])
    # overlap
    zo, zg = np.vdot(va, vb), cb.overlap(to(a), to(b), site.charges, ctx=ctx)
    if abs(zo - zg) > 1e-12 * na * nb:
        return "FAIL", tag + ": overlap %r vs %r" % (zo, zg)
    # entropies of the normalised state (CalcEE, src/utility.cpp:5-27: p = sigma^2, only p > 1e-12 counted, natural log)
    an = a.copy()
    an.A[0] = an.A[0] / na
    d = len(site.charges)
    eo = []
    for cut in range(1, n):
        p = np.linalg.svd((va / na).reshape(d ** cut, d ** (n - cut)), compute_uv=False) ** 2
        p = p[p > 1e-12]
        eo.append(float(-(p * np.log(p)).sum()))
    eg = np.array(cb.calc_ee(to(an), site.charges, ctx=ctx))
    if np.abs(np.array(eo) - eg).max() > 1e-10:
        return "FAIL", tag + ": entropies differ by %.3e" % np.abs(np.array(eo) - eg).max()
    # translation, exact (cutoff 1e-24): one of the two cyclic shifts of the state vector, the same one in every draw
    if n >= 3:
        t_g = cb.translate(common.to_cb(an), site.charges, 1e-24, ctx=ctx)
        vt = full(common.to_oracle(t_g, site))
        errs = [np.linalg.norm(vt - np.moveaxis(va / na, src, dst)) for src, dst in ((0, n - 1), (n - 1, 0))]
        which = int(np.argmin(errs))
        if min(errs) > 1e-9 or SHIFT.setdefault("dir", which) != which:
            return "FAIL", tag + ": translated state is not a cyclic shift (errors %.3e / %.3e, direction %d)" % (errs[0], errs[1], SHIFT["dir"])
    # <a|O|b> with a random dense MPO (nq = 1 entry point)
    if site.nq == 1:
        k = [1] + [int(rng.integers(1, 5)) for _ in range(n - 1)] + [1]
        d = len(site.charges)
        O = [rng.standard_normal((k[j], d, d, k[j + 1])) + (1j * rng.standard_normal((k[j], d, d, k[j + 1])) if cplx else 0) for j in range(n)]
        E = np.ones((1, 1, 1), dtype=complex)
        for j in range(n):      # E[a', k, b'] = sum E[a,k0,b] conj(A[a,t,a']) O[k0,t,s,k] B[b,s,b']
            E = np.einsum("akb,ate,ktsl,bsf->elf", E, a.A[j].conj(), O[j], b.A[j])
        want = E[0, 0, 0]
        got = cb.mpo_element(common.to_cb(a), cb.MPO(O, None, None, nq=1), common.to_cb(b), ctx=ctx)
        scale = na * nb * np.prod([np.linalg.norm(o) for o in O])
        if abs(want - got) > 1e-12 * scale:
            return "FAIL", tag + ": <a|O|b> %r vs %r" % (want, got)
    return "ok", ""


def main():
    import chain_b200 as cb
    from conftest import EMU_LIB
    ndraw = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ctx = cb.Context(lib_path=os.environ.get("CB200_FUZZ_LIB", EMU_LIB))      # (tools/asan_engine.py's instrumented build, for a run under ASan)
    nfail = nedge = 0
    t0 = time.time()
    for i in range(ndraw):
        try:
            status, msg = run_one(ctx, seed0 + i)
        except Exception as e:   # noqa: BLE001
            status, msg = "FAIL", "%s: %s" % (type(e).__name__, str(e)[:300])
        nfail += status == "FAIL"
        nedge += status == "edge"
        if status == "FAIL":
            print("FAIL seed %d: %s" % (seed0 + i, msg), flush=True)
    print("draws: %d  failures: %d  edge: %d  (%.0f s)" % (ndraw, nfail, nedge, time.time() - t0))
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
