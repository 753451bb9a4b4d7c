"""Row-sharded H_eff matvec, host logic on the CPU: the engine's shard planner (local rows of L, chunked
R-stage stores, penalty ownership) run on the serial emulation backend.  (a) virtual ranks in one process, (b) two
real processes exchanging their partial vectors over torch.distributed's gloo backend -- the exchange-free shard
mode (0) plus a host all-reduce stands in for the NCCL / peer-store exchange, which only exists on the GPU box
(tests/test_gpu_shard.py)."""
import os
import sys

import numpy as np
import pytest

import chain_b200 as cb
import common
from conftest import EMU_LIB, ROOT


CASES = [common.BOND_CASES[0], common.BOND_CASES[2], common.BOND_CASES[3], common.BOND_CASES[4]]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-%d-q%d" % (c[0], c[1], c[2], c[4]))
@pytest.mark.parametrize("world", [2, 3, 8])
def test_virtual_ranks_sum_to_full_matvec(emu_ctx, case, world):
    st, bc, n, cp, q, bond = case
    site, W, qk, c = common.capture_bond(st, bc, n, cp, q, bond)
    full = common.make_bond(emu_ctx, site, qk, c)
    rng = np.random.default_rng(5)
    o = rng.standard_normal(c["phi"].shape) * (np.abs(c["phi"]) > 0)
    full.set_penalty([o], 10.0)
    want = full.apply(c["phi"])
    flops = full.matvec_flops
    full.close()
    total = np.zeros_like(want)
    owned = np.zeros(want.shape, dtype=int)
    for rank in range(world):
        ctx = cb.Context(lib_path=EMU_LIB)
        ctx.shard(rank, world, mode="none", min_dim=0)
        B = common.make_bond(ctx, site, qk, c)
        B.set_penalty([o], 10.0)
        part = B.apply(c["phi"])
        assert abs(B.matvec_flops - flops) <= 1e-9 * flops or min(c["phi"].shape[0], 10 ** 9) < world * site.nq
        B.set_penalty([], 0.0)
        owned += (np.abs(B.apply(c["phi"])) > 0)
        total += part
        B.close()
        ctx.close()
    assert np.abs(total - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    assert owned.max() <= 1          # every output row is produced by exactly one rank


def test_small_bonds_stay_on_one_gpu(emu_ctx):
    # north_star: only D >= 1600 is partitioned; with the default threshold a small bond is computed in full by every rank
    st, bc, n, cp, q, bond = common.BOND_CASES[0]
    site, W, qk, c = common.capture_bond(st, bc, n, cp, q, bond)
    full = common.make_bond(emu_ctx, site, qk, c)
    want = full.apply(c["phi"])
    full.close()
    ctx = cb.Context(lib_path=EMU_LIB)
    ctx.shard(1, 2, mode="none")
    B = common.make_bond(ctx, site, qk, c)
    assert np.array_equal(B.apply(c["phi"]), want)
    B.close()


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st, bc, n, cp, ch, bond = common.BOND_CASES[4]            # complex128 Z3 block-sparse
        site, W, qk, c = common.capture_bond(st, bc, n, cp, ch, bond)
        ctx = cb.Context(lib_path=EMU_LIB)
        ctx.shard(rank, world, mode="none", min_dim=0)
        B = common.make_bond(ctx, site, qk, c)
        part = np.ascontiguousarray(B.apply(c["phi"]))
        t = torch.from_numpy(part.view(np.float64))
        dist.all_reduce(t)                                          # sum of zero-padded partial products
        got = t.numpy().view(part.dtype).reshape(part.shape)
        ref_ctx = cb.Context(lib_path=EMU_LIB)
        want = common.make_bond(ref_ctx, site, qk, c).apply(c["phi"])
        q.put((rank, float(np.abs(got - want).max()), float(np.abs(want).max())))
    finally:
        dist.destroy_process_group()


def test_two_processes_gloo_allreduce():
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [mpc.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, scale in res:
        assert err <= 1e-12 * max(1.0, scale)


def test_eigensolver_refuses_exchange_free_sharding(emu_ctx):
    """with mode 'none' Bond::apply returns partial rows (the caller sums the ranks); davidson / dmrg on such a context
    would compute Rayleigh quotients from partial vectors, so they must fail loudly instead."""
    st, bc, n, cp, q, bond = common.BOND_CASES[0]
    site, W, qk, c = common.capture_bond(st, bc, n, cp, q, bond)
    ctx = cb.Context(lib_path=EMU_LIB)
    ctx.shard(0, 2, mode="none", min_dim=0)
    B = common.make_bond(ctx, site, qk, c)
    with pytest.raises(cb.ChainB200Error):
        B.davidson(c["phi"], 2)
    B.close()
    site, W, qk = common.build_model("golden", "p", 6, [-1.0])
    H = cb.MPO(W, qk, site.charges, nq=1)
    psi0 = common.to_cb(common.warm_start(site, W, 6, 0, seed=3))
    with pytest.raises(cb.ChainB200Error):
        cb.dmrg(H, [], psi0, cb.Sweeps(1, 20, 1e-8), ctx=ctx)


def test_two_processes_p2p_exchange_on_host_emulation():
    """The exchange-dependent host logic between two REAL processes on the CPU: the body of tests/test_gpu_shard.py's worker (fused
    all-gather of the R-stage, slice-wise end-to-end call, Davidson, truncation spread over the ranks / sliced by columns, sharded DMRG
    runs in lockstep) on the host-emulation library, whose p2p landing areas are POSIX shared-memory segments with the CUDA backend's
    barrier protocol; rendezvous over torch.distributed's gloo backend."""
    import test_gpu_shard
    test_gpu_shard._run("p2p", lib=EMU_LIB, timeout=900)
