"""Differential fuzz of ONE bond update on random block structures (host logic of the engine on the emulation backend against the oracle):
the generator and the truncation / environment checks of tests/test_random_structures.py over many more draws, plus what the committed
tests do not combine -- COMPLEX penalty vectors (1-3 previous states, random weight) in the matvec, and the row-shard planner with 2-8 virtual ranks on ragged sectors (sectors shorter than the number of ranks, empty sectors): the partial
products of the ranks must sum to the unsharded product and every output row must come from exactly one rank.

    python tools/fuzz_bond.py [NDRAWS=200] [SEED0=0]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(ctx, seed):
    import chain_b200 as cb
    import test_random_structures as trs
    from conftest import EMU_LIB
    from oracle import dmrg as od
    rng = np.random.default_rng(70000 + seed)
    st = ["golden", "haagerup", "haagerupq"][rng.integers(3)]
    nq = 3 if st == "haagerupq" else 1
    cp = [float(np.round(rng.uniform(-1.5, 1.5), 3)) if rng.random() < 0.6 else 0.0 for _ in range(1 if st == "golden" else 3)]
    if all(c == 0 for c in cp):
        cp[0] = -1.0
    hi = 9 if st == "golden" else 6
    dims_l, dims_r = rng.integers(0, hi, nq), rng.integers(0, hi, nq)
    if dims_l.sum() == 0:
        dims_l[rng.integers(nq)] = 2
    if dims_r.sum() == 0:
        dims_r[rng.integers(nq)] = 2
    b = trs.random_bond(rng, st, cp, dims_l.tolist(), dims_r.tolist(), cplx_data=bool(rng.random() < 0.5))
    tag = "%s cp=%s Dl=%s Dr=%s cplx=%s" % (st, cp, dims_l.tolist(), dims_r.tolist(), np.iscomplexobj(b["phi"]))
    if not np.any(b["phi"]):
        return "edge", tag + ": no allowed block"
    try:
        trs.check(ctx, b)
    except AssertionError as e:
        return "FAIL", tag + ": truncate / env_update / apply check: %s" % str(e)[:200]
    site = b["site"]
    mk = lambda c: cb.Bond(b["L"], b["W1"], b["W2"], b["R"], ql=b["ql"], qr=b["qr"], qkl=b["kl"], qkm=b["km"], qkr=b["kr"],
                           site_charges=site.charges, nq=site.nq, ctx=c)
    cplx = np.iscomplexobj(b["phi"])
    # penalised operator with complex vectors
    P = int(rng.integers(1, 4))
    os_ = []
    for _ in range(P):
        o = (rng.standard_normal(b["phi"].shape) + (1j * rng.standard_normal(b["phi"].shape) if cplx else 0)) * b["pm"]
        os_.append(o / np.linalg.norm(o))
    w = float(rng.choice([1.0, 10.0, 1000.0]))
    B = mk(ctx)
    B.set_penalty(os_, w)
    phi = b["phi"]
    mv = lambda x: od.heff_apply(b["L"], b["W1"], b["W2"], b["R"], x) + sum(w * np.vdot(o, x) * o for o in os_)
    y, yo = B.apply(phi), mv(phi)
    if np.abs(y - yo).max() > 1e-11 * max(1.0, np.abs(yo).max()):
        B.close()
        return "FAIL", tag + ": penalised apply differs by %.3e (P=%d, w=%g)" % (np.abs(y - yo).max(), P, w)
    want = y
    B.close()
    # virtual ranks
    world = int(rng.integers(2, 9))
    total, owned = np.zeros_like(want), np.zeros(want.shape, dtype=int)
    for rank in range(world):
        c2 = cb.Context(lib_path=os.environ.get("CB200_FUZZ_LIB", EMU_LIB))
        c2.shard(rank, world, mode="none", min_dim=0)
        Bs = mk(c2)
        Bs.set_penalty(os_, w)
        total += Bs.apply(phi)
        Bs.set_penalty([], 0.0)
        owned += (np.abs(Bs.apply(phi)) > 0)
        Bs.close()
        c2.close()
    if np.abs(total - want).max() > 1e-11 * max(1.0, np.abs(want).max()) or owned.max() > 1:
        return "FAIL", tag + ": %d virtual ranks: sum differs by %.3e, max owners per element %d" % (world, np.abs(total - want).max(), owned.max())
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
