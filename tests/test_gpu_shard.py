"""-m gpu: the row-sharded H_eff matvec with its real exchanges, two processes.
  * p2p : all-gather fused into the R-stage epilogue through CUDA-IPC peer mappings + the flag barrier.  Runs on ONE GPU as
          well (both ranks on cuda:0, time-sliced), so the plumbing is covered on the single-GPU box;
  * nccl: ncclAllReduce of the zero-padded partials; needs two GPUs (NCCL refuses two ranks on one device).
Both are compared with the unsharded product on the same inputs, then a short sharded DMRG run with the unsharded one."""
import os
import sys

import numpy as np
import pytest

from conftest import PRODUCT_LIB, ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, mode, q, lib=PRODUCT_LIB):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import chain_b200 as cb
    import common
    from oracle import dmrg as od
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = rank % max(torch.cuda.device_count(), 1)          # (host emulation, tests/test_shard.py: no device)
        out = {}
        for case in (common.BOND_CASES[2], common.BOND_CASES[4]):           # dense real, complex Z3 block-sparse
            st, bc, n, cp, ch, bond = case
            site, W, qk, c = common.capture_bond(st, bc, n, cp, ch, bond)
            ref = cb.Context(device=dev, lib_path=lib)
            want = common.make_bond(ref, site, qk, c).apply(c["phi"])
            ctx = cb.Context(device=dev, lib_path=lib)
            ctx.shard(rank, world, mode=mode, landing_bytes=16 * c["phi"].size, min_dim=0)
            B = common.make_bond(ctx, site, qk, c)
            for rep in range(3):                                              # both landing halves + reuse
                got = B.apply(c["phi"] * (rep + 1))
                out["%s rep%d" % (st, rep)] = (float(np.abs(got - (rep + 1) * want).max()), float(np.abs(want).max()))
            # end-to-end form: each rank carries only its slice of the block-packed vector over PCIe (all-gather on the device side)
            xb = B.pack(c["phi"])
            b0, cnt = B.phi_blocks_slice()
            wantb = B.pack(want)
            for rep in range(2):
                ys = B.apply_blocks_sharded(xb[b0:b0 + cnt] * (rep + 1))
                out["%s slice rep%d" % (st, rep)] = (float(np.abs(ys - (rep + 1) * wantb[b0:b0 + cnt]).max()) if cnt else 0.0, float(np.abs(want).max()))
            ev, x, nmv = B.davidson(c["phi"], 2)
            evr, xr, _ = common.make_bond(ref, site, qk, c).davidson(c["phi"], 2)
            out["%s davidson" % st] = (abs(ev - evr), abs(evr))
            # svdBond: with Z_N blocks the eigenproblems are spread over the ranks and broadcast (nq > 1), else replicated
            Br = common.make_bond(ref, site, qk, c)
            for d in ("left", "right"):
                A1, B1, qm1, sp1, te1 = B.truncate(xr, d, 40, 1e-8)
                A2, B2, qm2, sp2, te2 = Br.truncate(xr, d, 40, 1e-8)
                out["%s trunc %s m" % (st, d)] = (float(abs(A1.shape[2] - A2.shape[2])), 0.0)
                out["%s trunc %s spec" % (st, d)] = (float(np.abs(np.sort(sp1) - np.sort(sp2)).max()), 1.0)
                r1 = np.tensordot(A1, B1, axes=([2], [0]))
                r2 = np.tensordot(A2, B2, axes=([2], [0]))
                out["%s trunc %s rec" % (st, d)] = (float(np.abs(r1 - r2).max()) * 1e-3, 1.0)     # eigenvectors near the cut: 1e-8 abs
            B.close()
        # dense truncation with its column-independent pieces sliced over the ranks (rho, the Q2 / Q1 back-transformation, X V): a dense
        # bond large enough for the two-stage eigensolver once its threshold is lowered, against the unsharded context
        os.environ["CB200_TWOSTAGE_MIN_N"] = "256"
        rng = np.random.default_rng(77)
        D, d, k = 96, 4, 3
        z = np.zeros(D, np.int32)
        for cplx in (False, True):
            def rnd(*shape):
                return rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
            E, Wd = np.asfortranarray(rnd(D, k, D)), np.asfortranarray(rnd(k, d, d, k))
            X = rnd(D * d, D * d) * (10.0 ** (-0.02 * np.arange(D * d)))[None, :]
            phi = np.asfortranarray((X / np.linalg.norm(X)).reshape(D, d, d, D, order="F"))
            kw = dict(ql=z, qr=z, qkl=np.zeros(k, np.int32), qkm=np.zeros(k, np.int32), qkr=np.zeros(k, np.int32), site_charges=np.zeros(d, np.int32), nq=1)
            ctx2 = cb.Context(device=dev, lib_path=lib)               # landing area sized for THIS phi
            ctx2.shard(rank, world, mode=mode, landing_bytes=16 * phi.size, min_dim=0)
            Bs = cb.Bond(E, Wd, Wd, E, ctx=ctx2, **kw)
            Bu = cb.Bond(E, Wd, Wd, E, ctx=ref, **kw)
            for dirn in ("left", "right"):
                A1, B1, _, sp1, te1 = Bs.truncate(phi, dirn, 150, 1e-12)
                A2, B2, _, sp2, te2 = Bu.truncate(phi, dirn, 150, 1e-12)
                tag = "dense two-stage %s %s" % ("c128" if cplx else "f64", dirn)
                out[tag + " m"] = (float(abs(len(sp1) - len(sp2))), 0.0)
                out[tag + " spec"] = (float(np.abs(np.sort(sp1) - np.sort(sp2)).max()), 0.01)
                out[tag + " rec"] = (float(np.abs(np.tensordot(A1, B1, axes=([2], [0])) - np.tensordot(A2, B2, axes=([2], [0]))).max()) * 1e-2, 1.0)
            Bs.close(); Bu.close()
        del os.environ["CB200_TWOSTAGE_MIN_N"]
        # a short sharded DMRG run (every rank drives the same sweep in lockstep; min_dim=0 shards every bond)
        site, W, qk = common.build_model("golden", "p", 6, [-1.0])
        H = cb.MPO(W, qk, site.charges, nq=1)
        psi0 = common.to_cb(od.product_mps(site, 6, 0, np.random.default_rng(0)))
        sw = cb.Sweeps(4, [10, 20, 40, 80], [1e-5, 1e-6, 1e-7, 1e-8])
        e1, _ = cb.dmrg(H, [], psi0, sw, ctx=ref)
        e2, _ = cb.dmrg(H, [], psi0, sw, ctx=ctx)
        out["dmrg"] = (abs(e1 - e2), abs(e1))
        # the same with Z3 blocks (complex128): sharded matvec + spread truncation + replicated rest must track the 1-GPU run
        site, W, qk = common.build_model("haagerupq", "p", 6, [0.0, -1.0, 0.0])
        H = cb.MPO(W, qk, site.charges, nq=site.nq)
        psi0 = common.to_cb(common.warm_start(site, W, 6, 0, seed=7))
        sw = cb.Sweeps(3, [10, 20, 40], [1e-5, 1e-6, 1e-7])
        e1, _ = cb.dmrg(H, [], psi0, sw, ctx=ref)
        e2, _ = cb.dmrg(H, [], psi0, sw, ctx=ctx)
        out["dmrg z3"] = (abs(e1 - e2), abs(e1) * 10)
        q.put((rank, out))
    except BaseException:                      # report instead of leaving the parent to wait for its queue timeout
        import traceback
        q.put((rank, {"worker exception": traceback.format_exc()}))
        raise
    finally:
        dist.destroy_process_group()


def _run(mode, lib=PRODUCT_LIB, timeout=300):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [mpc.Process(target=_worker, args=(r, 2, port, mode, q, lib)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=timeout) for _ in procs]
    for rank, out in res:
        assert "worker exception" not in out, "rank %d:\n%s" % (rank, out.get("worker exception"))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        for k, (err, scale) in out.items():
            assert err <= 1e-11 * max(1.0, scale), (rank, k, err, scale)


def test_p2p_fused_allgather_two_ranks():
    _run("p2p")


def test_nccl_allreduce_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("ncclAllReduce needs two GPUs (two ranks cannot share one device)")
    _run("nccl")


def _bench_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import chain_b200 as cb
    import bench
    import test_gpu_bench_parity as tb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = bench.build_workload("cfg2", seed=2000)
        ctx = cb.Context(device=rank, lib_path=PRODUCT_LIB)
        ctx.shard(rank, world, mode="p2p", landing_bytes=16 * w["phi"].size, min_dim=0)
        rows = np.array([0, 266, 267, 533, 534, 1066, 1599])           # both ranks' row slices of every sector
        rel = tb._check_rows(ctx, w, rows, 1.0)
        q.put((rank, {"cfg2 rows": (rel, 1e11)}))                      # _check_rows asserted the bound already
    except BaseException:
        import traceback
        q.put((rank, {"worker exception": traceback.format_exc()}))
        raise
    finally:
        dist.destroy_process_group()


def test_cfg2_bench_workload_two_ranks_rows_match_numpy():
    """The benchmarked D = 1600 matvec row-sharded over two GPUs (fused all-gather), sampled rows against NumPy."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29500 + (os.getpid() + 7) % 2000
    procs = [mpc.Process(target=_bench_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for rank, out in res:
        assert "worker exception" not in out, "rank %d:\n%s" % (rank, out.get("worker exception"))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(res) == 2
