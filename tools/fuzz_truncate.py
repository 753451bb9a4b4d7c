"""Robustness sweep of MPS::svdBond (engine split_bond) on the HOST-EMULATION build against the oracle: random Z3 / dense structures (ragged and
empty sectors) whose two-site wavefunction is BUILT from a designed Schmidt spectrum -- exact multiplets, pairs split at rounding level, a
graded tail down to 1e-18, rank-deficient blocks -- and truncated with maxdim landing inside a multiplet, tight and loose cutoffs, both
directions.  Checks what the parity tests check: kept dimension per sector, spectrum, truncation error, isometry, the two-site tensor A.B.

    python tools/fuzz_truncate.py [nseeds]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import chain_b200 as cb  # noqa: E402
from oracle import dmrg as od  # noqa: E402

EMU = os.path.join(ROOT, "tests", "hostemu", "libcb200_hostemu.so")


def block_isometry(rng, rowq, colq, nq, cplx):
    """M[row, col] with orthonormal columns, nonzero only where rowq == colq; None if some charge has more columns than rows."""
    M = np.zeros((len(rowq), len(colq)), dtype=complex if cplx else float)
    for q in range(nq):
        r, c = np.nonzero(rowq == q)[0], np.nonzero(colq == q)[0]
        if len(c) == 0:
            continue
        if len(c) > len(r):
            return None
        X = rng.standard_normal((len(r), len(c))) + (1j * rng.standard_normal((len(r), len(c))) if cplx else 0)
        Q, _ = np.linalg.qr(X)
        M[np.ix_(r, c)] = Q
    return M


def spectrum(rng, m, kind):
    if kind == 0:      # geometric decay with exact multiplets
        s = 10.0 ** (-rng.uniform(0.05, 0.6) * np.arange(m))
        for _ in range(3):
            i = int(rng.integers(0, max(m - 3, 1)))
            s[i:i + int(rng.integers(2, 4))] = s[i]
    elif kind == 1:    # pairs split at rounding level + a graded tail
        s = 10.0 ** (-9.0 * np.arange(m) / max(m - 1, 1))
        s[1::2] = s[0::2][: len(s[1::2])] * (1 + 4e-16)
    elif kind == 2:    # flat (maximally degenerate)
        s = np.ones(m)
    else:              # few large + many exactly zero (rank-deficient rho)
        s = np.zeros(m)
        k = max(1, m // 4)
        s[:k] = rng.random(k) + 0.1
    return s / np.linalg.norm(s)


def one_case(ctx, rng, nq, cplx, worst, fails, tag):
    d = 6 if nq == 3 else int(rng.integers(2, 5))
    sq = np.array([0, 1, 2, 0, 1, 2], np.int32) if nq == 3 else np.zeros(d, np.int32)
    hi = 9 if rng.random() < 0.75 else 41          # a quarter of the draws are large enough for the compact-WY back-transformation
    Dl, Dr, Dm = (int(rng.integers(1, hi)) for _ in range(3))
    Dm = max(1, min(Dm * 3, Dl * d // 2, Dr * d // 2))
    ql = rng.integers(0, nq, Dl).astype(np.int32)
    qr = rng.integers(0, nq, Dr).astype(np.int32)
    qm = rng.integers(0, nq, Dm).astype(np.int32)
    rowq = ((ql[:, None] + sq[None, :]) % nq).reshape(-1)              # (l, s1), s1 fastest in C order -> matches phi.reshape(Dl*d, .)
    colq = ((qr[None, :] - sq[:, None]) % nq).reshape(-1)              # (s2, r)
    U = block_isometry(rng, rowq, qm, nq, cplx)
    V = block_isometry(rng, colq, qm, nq, cplx)
    if U is None or V is None:
        return
    sig = spectrum(rng, Dm, int(rng.integers(0, 4)))
    phi = ((U * sig) @ V.conj().T).reshape(Dl, d, d, Dr)
    nrm = np.linalg.norm(phi)
    if nrm == 0:
        return
    phi = np.asfortranarray(phi / nrm)
    k = 2
    E = lambda D, q: (rng.standard_normal((D, k, D)) + (1j * rng.standard_normal((D, k, D)) if cplx else 0)) * (q[:, None, None] == q[None, None, :])
    W = np.zeros((k, d, d, k))
    for a in range(k):
        W[a, :, :, a] = np.eye(d)
    zk = np.zeros(k, np.int32)
    B = cb.Bond(np.asfortranarray(E(Dl, ql)), np.asfortranarray(W), np.asfortranarray(W), np.asfortranarray(E(Dr, qr)), ql=ql, qr=qr,
                qkl=zk, qkm=zk, qkr=zk, site_charges=sq, nq=nq, ctx=ctx)
    try:
        for direction in ("left", "right"):
            # (cutoffs off the powers of ten: the designed spectra are, and `truncerr + P(n) < cutoff` must not sit on a knife edge)
            for maxdim, cutoff in ((10 ** 6, 3e-13), (max(1, Dm // 2), 1e-12), (max(1, Dm - 1), 3e-8), (10 ** 6, 3e-4)):
                key = (tag, direction)
                try:
                    A1, B1, qg, spec, terr = B.truncate(phi, direction, maxdim, cutoff)
                except Exception as e:   # noqa: BLE001
                    fails.append((key, maxdim, cutoff, str(e)))
                    continue
                A2, B2, qo, speco, terro = od.split_bond(phi, ql, qr, sq, nq, direction, maxdim, 1, cutoff)
                r = [0.0] * 5
                if len(spec) != len(speco) or not np.array_equal(np.bincount(qg, minlength=nq), np.bincount(qo, minlength=nq)):
                    # a cut between two values that differ only in their last bits may legitimately fall either side
                    so = np.sort(np.concatenate([speco, np.zeros(1)]))[::-1]
                    edge = abs(len(spec) - len(speco)) <= 2 and terr <= max(cutoff, terro) * (1 + 1e-9) + 1e-15
                    if not edge:
                        fails.append((key, maxdim, cutoff, "kept %d vs oracle %d" % (len(spec), len(speco)), so[:3].tolist()))
                    continue
                r[0] = np.abs(np.sort(spec) - np.sort(speco)).max() / 1e-12
                r[1] = abs(terr - terro) / 1e-14
                iso = A1 if direction == "left" else B1
                m = len(spec)
                g = np.tensordot(iso.conj(), iso, axes=([0, 1], [0, 1])) if direction == "left" else np.tensordot(iso, iso.conj(), axes=([1, 2], [1, 2]))
                r[2] = np.abs(g - np.eye(m)).max() / 1e-12
                ab, abo = np.tensordot(A1, B1, axes=([2], [0])), np.tensordot(A2, B2, axes=([2], [0]))
                # the two-site tensor itself is only defined when the cut does not fall inside a (near-)multiplet; the fidelity with phi always is
                pooled = np.sort(np.linalg.svd(phi.reshape(Dl * d, d * Dr), compute_uv=False) ** 2)[::-1]
                nxt = pooled[m] if m < len(pooled) else 0.0
                if pooled[m - 1] - nxt > 1e-6 * pooled[m - 1]:
                    r[3] = np.abs(ab - abo).max() / 1e-9
                else:
                    r[3] = abs(abs(np.vdot(ab, phi)) - abs(np.vdot(abo, phi))) / 1e-12
                r[4] = abs(np.linalg.norm(ab) - 1) / 1e-12
                cur = worst.setdefault(key, [0.0] * 5)
                worst[key] = [max(a, b) for a, b in zip(cur, r)]
                if max(r) >= 1:
                    fails.append((key, maxdim, cutoff, "ratio", [round(x, 2) for x in r], (Dl, Dr, Dm, d)))
    finally:
        B.close()


def main():
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    ctx = cb.Context(lib_path=EMU)
    worst, fails = {}, []
    for seed in range(nseeds):
        for nq in (3, 1):
            for cplx in (False, True):
                rng = np.random.default_rng(7000 + 10 * seed + nq + cplx)
                one_case(ctx, rng, nq, cplx, worst, fails, "%s %s" % ("Z3" if nq == 3 else "dense", "c128" if cplx else "f64"))
    print("ratio measured/tolerance: spectrum 1e-12 | truncerr 1e-14 | isometry 1e-12 | A.B vs oracle 1e-9 | norm 1e-12")
    for k, v in sorted(worst.items()):
        print("%-12s %-6s " % k + "  ".join("%.3f" % x for x in v))
    print("failures: %d" % len(fails))
    for f in fails[:30]:
        print("  ", f)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
