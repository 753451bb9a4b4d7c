"""Robustness sweep of the truncation's eigensolver driver on the HOST-EMULATION build (no GPU): the engine logic -- shifts, inverse iteration
start vectors, Newton-Schulz orthonormalisation, one-/two-stage plumbing, back-transformation -- is the product's own engine.cpp; only the
kernels are the serial restatements of tests/hostemu.  Random seeds x spectrum families that stress it (geometric decay into rounding noise,
exactly repeated columns, exact multiplets, a shifted copy whose whole tail is one numerically degenerate cluster), all eigenvectors requested.
Prints the worst ratio measured / tolerance per check; anything >= 1 would be a failing GPU/CPU test on some input.

    python tools/fuzz_heig.py [nseeds]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chain_b200 as cb  # noqa: E402

EMU = os.path.join(ROOT, "tests", "hostemu", "libcb200_hostemu.so")


def families(rng, n, cplx):
    def rnd(*s):
        return rng.standard_normal(s) + (1j * rng.standard_normal(s) if cplx else 0)
    X = rnd(n + 5, n) * (10.0 ** (-0.04 * np.arange(n)))
    X[:, 7] = X[:, 3]
    A = X.conj().T @ X
    A /= np.trace(A).real
    yield "decay+repeat", A
    yield "decay+repeat shifted (cluster of ~n/2)", A + 0.01 * np.eye(n)
    Q, _ = np.linalg.qr(rnd(n, n))
    w = np.sort(rng.random(n))[::-1]
    w[5:9] = w[5]                      # exact 4-fold multiplet
    w[20:22] = w[20] * (1 + 1e-15)     # a pair split at rounding level
    M = (Q * w) @ Q.conj().T
    yield "multiplets", (M + M.conj().T) / 2
    w2 = 10.0 ** (-16.0 * np.arange(n) / n)
    M2 = (Q * w2) @ Q.conj().T
    yield "graded 1..1e-16", (M2 + M2.conj().T) / 2


def main():
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    ctx = cb.Context(lib_path=EMU)
    worst = {}
    fails = []
    for two_stage in (False, True):
        os.environ["CB200_TWOSTAGE_MIN_N"] = "256" if two_stage else "0"
        for cplx in (False, True):
            for n in (257, 300):
                for seed in range(nseeds):
                    rng = np.random.default_rng(1000 * seed + n + cplx)
                    for name, A in families(rng, n, cplx):
                        for m in (n, n // 3):
                            key = (name, "two-stage" if two_stage else "one-stage")
                            try:
                                w, V, it = cb.k_heig(A, m, ctx=ctx)
                            except Exception as e:   # noqa: BLE001
                                fails.append((key, cplx, n, m, seed, str(e)))
                                continue
                            ev = np.linalg.eigvalsh(A)[::-1]
                            sc = max(abs(ev).max(), 1e-300)
                            r = [np.abs(w - ev).max() / sc / (1e-14 * np.sqrt(n)),
                                 np.abs(V.conj().T @ V - np.eye(m)).max() / 1e-13,
                                 np.abs(A @ V - V @ (V.conj().T @ A @ V)).max() / sc / (2e-14 * np.sqrt(n)),
                                 it / 400.0]
                            cur = worst.setdefault(key, [0, 0, 0, 0])
                            worst[key] = [max(a, b) for a, b in zip(cur, r)]
    print("ratio measured/tolerance: eigenvalues | orthonormality | invariance residual | NS passes/cap")
    for k, v in sorted(worst.items()):
        print("%-45s %-10s  %.2f  %.2f  %.2f  %.3f" % (k[0], k[1], *v))
    print("failures: %d" % len(fails))
    for f in fails[:20]:
        print("  ", f)
    return 1 if fails or any(max(v[:3]) >= 1 for v in worst.values()) else 0


if __name__ == "__main__":
    sys.exit(main())
