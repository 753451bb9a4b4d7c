"""Differential fuzz of the multi-rank engine paths on the host emulation: `world` real processes (gloo rendezvous, the p2p exchange over POSIX
shared memory with the same landing / barrier protocol as the CUDA backend) run the random `dmrg()` problems of tools/fuzz_dmrg.py with EVERY bond
sharded (min_dim = 0) -- row-sharded matvec with the all-gather in the R-stage epilogue, charge blocks of the truncation spread over the ranks
with broadcasts, column-sliced dense truncation, replicated rest -- and compare, bond by bond, with the SAME engine unsharded in the same process.
Also asserted: every rank ends with bit-identical energies (the replicated parts must stay in lockstep or the ranks' sweeps diverge).

    python tools/fuzz_shard.py [NDRAWS=40] [WORLD=2] [SEED0=0]
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def worker(rank, world, port, ndraw, seed0, q):
    import torch.distributed as dist
    import chain_b200 as cb
    import common
    import fuzz_dmrg
    from conftest import EMU_LIB
    from oracle import dmrg as od
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = dict(draws=0, fail=[], worst=0.0, flips=0, launches_ref=0, launches_sharded=0)
    try:
        lib = os.environ.get("CB200_FUZZ_LIB", EMU_LIB)
        ref = cb.Context(lib_path=lib)
        ctx = cb.Context(lib_path=lib)
        ctx.shard(rank, world, mode="p2p", landing_bytes=64 << 20, min_dim=0)
        for i in range(ndraw):
            seed = seed0 + i
            p = fuzz_dmrg.draw(np.random.default_rng(1000 + seed))
            site, W, qk = common.build_model(p["st"], p["bc"], p["n"], p["cp"], U=p["U"])
            H = cb.MPO(W, qk, site.charges, nq=site.nq)
            cplx = any(np.iscomplexobj(w) and np.abs(np.imag(w)).max() > 0 for w in W)
            rng = np.random.default_rng(seed)
            sw = cb.Sweeps(p["nsweep"], p["maxd"], p["cut"], niter=p["niter"], mindim=p["mind"])
            i1, i2 = cb.DmrgInfo(), cb.DmrgInfo()
            status, done, e2, st1, st2 = "ok", [], [], [], []
            try:
                for s_ in range(p["nstates"]):           # the second state is penalised against the first (penalty rows are sharded too)
                    if p["product"]:
                        psi0 = od.product_mps(site, p["n"], p["q"], rng, dtype=complex if cplx else float)
                    else:
                        psi0 = common.warm_start(site, W, p["n"], p["q"], seed=seed + s_, maxdim=int(rng.choice([2, 4, 8])))
                    ea, pa = cb.dmrg(H, done, common.to_cb(psi0), sw, weight=p["weight"], ctx=ref, info=i1)
                    eb, pb = cb.dmrg(H, done, common.to_cb(psi0), sw, weight=p["weight"], ctx=ctx, info=i2)
                    e2.append(eb)
                    st1 += i1.stats
                    st2 += i2.stats
                    done.append(pa)
            except cb.ChainB200Error as e:      # the same refusal on every rank is fine (e.g. a schedule the engine rejects); anything else is not
                status = "error: %s" % str(e)[:200]
                e2, st2 = [float("nan")], []
            i1.stats, i2.stats = st1, st2
            # lockstep: every rank must hold the same sharded result, bit for bit
            digest = hashlib.sha1(repr((e2, [(s["m"], s["energy"], s["truncerr"]) for s in i2.stats])).encode()).hexdigest()
            allg = [None] * world
            dist.all_gather_object(allg, (status, digest))
            if len(set(allg)) != 1:
                out["fail"].append((seed, "ranks disagree: %s" % (allg,)))
                break                            # the ranks are out of step: nothing after this is meaningful
            if status != "ok":
                out["fail"].append((seed, status))
                continue
            out["draws"] += 1
            out["launches_ref"], out["launches_sharded"] = ref.launches, ctx.launches
            bad = None
            for a, b in zip(i1.stats, i2.stats):
                if a["m"] != b["m"]:
                    out["flips"] += 1            # knife-edge truncation decisions flip under re-ordered sums (tools/fuzz_dmrg.py's `edge`)
                    break
                de = abs(a["energy"] - b["energy"]) / max(1.0, abs(a["energy"]))
                out["worst"] = max(out["worst"], de)
                if de > 1e-8:
                    bad = "sweep %d bond %d: energy differs by %.3e from the unsharded run" % (a["sweep"], a["bond"], de)
                    break
            if bad:
                out["fail"].append((seed, "%s %s" % (p, bad)))
        q.put((rank, out))
    except BaseException:
        import traceback
        q.put((rank, {"worker exception": traceback.format_exc()}))
        raise
    finally:
        dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp
    ndraw = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    t0 = time.time()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [mpc.Process(target=worker, args=(r, world, port, ndraw, seed0, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=3000) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    nfail = 0
    for rank, out in sorted(res):
        if "worker exception" in out:
            print("rank %d crashed:\n%s" % (rank, out["worker exception"]))
            nfail += 1
            continue
        for seed, msg in out["fail"]:
            print("FAIL rank %d seed %d: %s" % (rank, seed, msg))
        nfail += len(out["fail"])
        print("rank %d: draws %d  failures %d  knife-edge flips %d  worst energy gap to the unsharded run %.2e  launches %d unsharded / %d sharded" % (
            rank, out["draws"], len(out["fail"]), out["flips"], out["worst"], out["launches_ref"], out["launches_sharded"]))
    print("world: %d  draws: %d  failures: %d  (%.0f s)" % (world, ndraw, nfail, time.time() - t0))
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
