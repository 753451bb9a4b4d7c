"""Differential fuzz of whole `dmrg()` calls: the engine (host-emulation build; same engine sources as the product) against the oracle on
randomly drawn problems, bond by bond.

The committed parity cases use the documented models and Chain's schedule (maxdim 10..200, cutoff 1e-5..1e-9, niter 2, mindim 1).  This draws
what those never reach: all three site types, every boundary condition of src/dmrg.h (p, o, s, sp), lengths 3..7, random couplings and U, every
total charge, `niter` 1..4, `mindim` 1..5 (also above maxdim), maxdim schedules that shrink as well as grow and go down to 1, cutoffs from 1e-3
to 1e-11 and 0, 1..3 sweeps, product-state and warm starts, dense and block-packed start states, and a penalised second state with a random weight.  Compared per bond update: the
kept dimension m, the Davidson eigenvalue and the truncation error; at the end energy, overlap of the states and the entanglement entropies.
These schedules leave the state far from converged, where a rounding-level difference is amplified from bond to bond (the oracle against ITSELF
from a start perturbed by 1e-15 drifts to 1e-6 within three sweeps at maxdim 2-3), so the yardstick is the oracle's own sensitivity: every draw
runs the oracle twice (exact; and from a start perturbed by 1e-15 with every H_eff product perturbed by 2e-16) and the engine may deviate from the oracle by at most a multiple of the largest deviation the
two oracle runs have shown up to that bond) -- max(1e-8, 1000 x ...) for energies and truncation errors (a sweep at cutoff 1e-13 keeps Schmidt weights of 1e-12, whose vectors are only defined to ~1e-10: the next truncating sweep shows that as 1e-9 in the energy); this fuzz is after gross disagreements (a wrong block, a stale
cache, a mishandled schedule), the committed parity tests hold the 1e-10 / 1e-8 tolerances on the documented models.  Also `edge`: a run in which
the oracle's davidson() had to replace a dependent residual by a random vector (the start already was an eigenvector): generator-dependent from
there on, in ITensor too.  A bond update that is ASKED to keep states of zero weight (mindim above the rank, cutoff 0 on a rank-deficient block: the
oracle's last kept weight is below 1e-11 of its largest -- a singular vector of relative weight w is defined to eps/sqrt(w) only, 3e-11 there --, or the kept dimensions differ in a sweep with cutoff 0 / mindim > 1) decides on the sign
of rounding noise, and one whose cut runs through a multiplet (relative gap at the cut below 1e-6) picks an arbitrary vector of it -- in ITensor
as well; such a draw is reported as `edge` from that bond on and not counted.

    python tools/fuzz_dmrg.py [NDRAWS=60] [SEED0=0]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def draw(rng):
    if os.environ.get("CB200_FUZZ_BIG") == "1":
        # fewer, larger problems: density-matrix blocks of order 250-500, where (with CB200_TWOSTAGE_MIN_N=256) the two-stage eigensolver and,
        # under tools/fuzz_shard.py, the column-sliced dense truncation take over
        st = ["haagerup", "haagerupq"][rng.integers(2)]
        cp = [float(np.round(rng.uniform(-1.5, 1.5), 3)) for _ in range(3)]
        return dict(st=st, bc=["p", "o"][rng.integers(2)], n=6, cp=cp, U=2.0, q=int(rng.integers(3)) if st == "haagerupq" else 0, nsweep=1,
                    maxd=[int(rng.choice([45, 60, 80]))], cut=[float(rng.choice([1e-8, 1e-10]))], niter=[2], mind=[1], product=False, nstates=1,
                    weight=1000.0)
    st = ["golden", "haagerup", "haagerupq"][rng.integers(3)]
    bc = ["p", "o", "s", "sp"][rng.integers(4)]
    n = int(rng.integers(3, 8 if st == "golden" else 6))
    ncp = 1 if st == "golden" else 3
    cp = [float(np.round(rng.uniform(-1.5, 1.5), 3)) if rng.random() < 0.7 else 0.0 for _ in range(ncp)]
    if all(c == 0 for c in cp):
        cp[0] = -1.0
    U = float(np.round(rng.uniform(0.5, 3.0), 2))
    q = int(rng.integers(3)) if st == "haagerupq" else 0
    nsweep = int(rng.integers(1, 4))
    maxd = [int(rng.choice([1, 2, 3, 5, 8, 13, 30, 100])) for _ in range(nsweep)]
    cut = [float(rng.choice([1e-3, 1e-5, 1e-7, 1e-9, 1e-11, 1e-13, 1e-13, 0.0])) for _ in range(nsweep)]
    niter = [int(rng.integers(1, 5)) for _ in range(nsweep)]
    mind = [int(rng.choice([1, 1, 1, 1, 1, 1, 2, 5])) for _ in range(nsweep)]
    return dict(st=st, bc=bc, n=n, cp=cp, U=U, q=q, nsweep=nsweep, maxd=maxd, cut=cut, niter=niter, mind=mind,
                product=bool(rng.random() < 0.4), nstates=int(rng.choice([1, 1, 2])), weight=float(rng.choice([1000.0, 10.0, 1.0])))


def run_one(ctx, p, seed):
    import chain_b200 as cb
    import common
    from oracle import dmrg as od
    site, W, qk = common.build_model(p["st"], p["bc"], p["n"], p["cp"], U=p["U"])
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    cplx = any(np.iscomplexobj(w) and np.abs(np.imag(w)).max() > 0 for w in W)
    rng = np.random.default_rng(seed)
    done_o, done_g = [], []
    worst = dict(energy=0.0, terr=0.0, ee=0.0, ov=0.0)
    for s in range(p["nstates"]):
        if p["product"]:
            psi0 = od.product_mps(site, p["n"], p["q"], rng, dtype=complex if cplx else float)
        else:
            psi0 = common.warm_start(site, W, p["n"], p["q"], seed=seed + s, maxdim=int(rng.choice([2, 4, 8])))
        so, so2 = [], []
        swo = od.Sweeps(p["nsweep"], p["maxd"], p["cut"], niter=p["niter"], mindim=p["mind"])
        swg = cb.Sweeps(p["nsweep"], p["maxd"], p["cut"], niter=p["niter"], mindim=p["mind"])
        nrand = od.RANDOMIZED[0]
        eo, po = od.dmrg(W, done_o, psi0, swo, weight=p["weight"], stats=so)
        if od.RANDOMIZED[0] != nrand:
            return "edge", "state %d: davidson replaced a dependent residual by a random vector (generator-dependent from there on, in ITensor too)" % s, worst
        psi1 = psi0.copy()                                   # the oracle's own sensitivity: the same run from a start perturbed by 1e-15
        prng = np.random.default_rng(77 + seed)
        psi1.A = [a * (1 + 1e-15 * prng.standard_normal(a.shape)) for a in psi1.A]
        # ... and with every H_eff product perturbed at rounding level (a product state absorbs a perturbation of the START in its
        # normalisation; what the engine differs in from the oracle is fresh rounding in every product)
        exact_apply = od.heff_apply
        od.heff_apply = lambda *a_: (lambda y: y * (1 + 2e-16 * prng.standard_normal(y.shape)))(exact_apply(*a_))
        try:
            eo2, po2 = od.dmrg(W, done_o, psi1, swo, weight=p["weight"], stats=so2)
        finally:
            od.heff_apply = exact_apply
        info = cb.DmrgInfo()
        packed = bool(np.random.default_rng(5 + seed + s).random() < 0.3)     # ITensor's block-packed QDense storage at the boundary (layout = 1)
        start = cb.PackedMPS.from_dense(common.to_cb(psi0), site.charges) if packed else common.to_cb(psi0)
        eg, pg = cb.dmrg(H, done_g, start, swg, weight=p["weight"], ctx=ctx, info=info)
        if packed:
            pg = pg.to_dense()
        if len(so) != len(info.stats):
            return "FAIL", "number of bond updates %d vs %d" % (len(so), len(info.stats)), worst
        self_e = self_t = 0.0
        for a, a2, b in zip(so, so2, info.stats):
            where = "state %d sweep %d bond %d" % (s, a["sweep"], a["bond"])
            if a["last_kept_rel"] < 1e-11 or a["cut_gap"] < 1e-6:
                return "edge", where + ": cut at a zero-weight state / through a multiplet (m %d vs %d)" % (a["m"], b["m"]), worst
            if a["m"] != a2["m"]:
                return "edge", where + ": the oracle's own kept dimension flips under a 1e-15 perturbation", worst
            self_e = max(self_e, abs(a["energy"] - a2["energy"]) / max(1.0, abs(a["energy"])))
            self_t = max(self_t, abs(a["truncerr"] - a2["truncerr"]))
            if a["m"] != b["m"]:
                zero_asked = p["cut"][a["sweep"]] == 0.0 or p["mind"][a["sweep"]] > 1
                return ("edge" if zero_asked else "FAIL"), where + ": m %d vs %d" % (a["m"], b["m"]), worst
            de = abs(a["energy"] - b["energy"]) / max(1.0, abs(a["energy"]))
            dt = abs(a["truncerr"] - b["truncerr"])
            worst["energy"] = max(worst["energy"], de / max(1e-8, 1000 * self_e))
            worst["terr"] = max(worst["terr"], dt / max(1e-8, 1000 * self_t))
            if de > max(1e-8, 1000 * self_e) or dt > max(1e-8, 1000 * self_t):
                return "FAIL", where + ": energy %.3e (oracle vs itself %.3e) truncerr %.3e (%.3e)" % (de, self_e, dt, self_t), worst
        pgo = common.to_oracle(pg, site)
        ov = abs(abs(od.overlap(po, pgo)) - 1)
        ee = float(np.abs(np.array(od.calc_ee(po)) - np.array(od.calc_ee(pgo))).max())
        ov2 = abs(abs(od.overlap(po, po2)) - 1)
        ee2 = float(np.abs(np.array(od.calc_ee(po)) - np.array(od.calc_ee(po2))).max())
        worst["ov"] = max(worst["ov"], ov / max(1e-8, 1000 * ov2))
        worst["ee"] = max(worst["ee"], ee / max(1e-7, 1000 * ee2))
        if abs(eo - eg) > max(1e-8, 1000 * abs(eo - eo2)) * max(1.0, abs(eo)) or ov > max(1e-8, 1000 * ov2) or ee > max(1e-7, 1000 * ee2):
            return "FAIL", "state %d result: energy %.3e (%.3e) overlap %.3e (%.3e) ee %.3e (%.3e)" % (s, abs(eo - eg), abs(eo - eo2), ov, ov2, ee, ee2), worst
        if abs(od.overlap(pgo, pgo) - 1) > 1e-10:
            return "FAIL", "state %d is not normalised" % s, worst
        done_o.append(po)
        done_g.append(common.to_cb(po))      # the next (penalised) state is penalised against the SAME previous state on both sides
    return "ok", "", worst


def main():
    import chain_b200 as cb
    from conftest import EMU_LIB
    ndraw = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ctx = cb.Context(lib_path=os.environ.get("CB200_FUZZ_LIB", EMU_LIB))      # (tools/asan_engine.py's instrumented build, for a run under ASan)
    nfail = nedge = 0
    worst = dict(energy=0.0, terr=0.0, ee=0.0, ov=0.0)
    t0 = time.time()
    for i in range(ndraw):
        seed = seed0 + i
        p = draw(np.random.default_rng(1000 + seed))
        try:
            status, msg, w = run_one(ctx, p, seed)
        except Exception as e:   # noqa: BLE001
            status, msg, w = "FAIL", "%s: %s" % (type(e).__name__, e), {}
        for k, v in w.items():
            worst[k] = max(worst[k], v)
        nfail += status == "FAIL"
        nedge += status == "edge"
        if status == "FAIL":
            print("%s seed %d %s\n     %s" % (status, seed, p, msg), flush=True)
    print("draws: %d  failures: %d  edge: %d  worst measured/allowed %s  (%.0f s)" % (ndraw, nfail, nedge, {k: float("%.2e" % v) for k, v in worst.items()}, time.time() - t0))
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
