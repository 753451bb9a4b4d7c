"""DMRG sweep wall-time at large bond dimension (BASELINE metric, first half): ONE full sweep (2(L-1) bond updates: 3 H_eff
matvecs + Davidson level-1 + truncation to maxdim + environment update each) through cb200_dmrg on a synthetic MPS.

The physical runs never reach D >= 800 at the documented cutoffs, so the state is a seeded random RIGHT-CANONICAL
block-sparse MPS with the capped bond profile min(d^b, d^(L-b), D) (every site tensor is the Q factor of an i.i.d.
N(0,1) matrix, per charge block) and the Hamiltonian is the model's true compressed MPO.  cutoff = 1e-12 so that maxdim
binds while svdBond stays on ITensor's denmatDecomp branch (cutoff >= 1e-12, the branch every documented run of the reference takes;
CB200_SWEEP_CUTOFF=1e-16 times the true-SVD branch instead).

  python tools/sweep_bench.py haagerupq p 12 1600 0,-1,0      # BASELINE configs[1] at a chain long enough to reach D
  python tools/sweep_bench.py golden p 24 800 -1               # BASELINE configs[2]
  CB200_SWEEP_PACKED=1 python tools/sweep_bench.py ...         # host MPS in ITensor's QDense block storage instead of dense arrays
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chain_b200 as cb  # noqa: E402
from chain_b200 import models  # noqa: E402


def link_profile(ss, n, D, qtot=0):
    nq, d = ss.nq, ss.d
    ds = np.bincount(ss.charges, minlength=nq)
    left = [np.zeros(nq, np.int64)]
    left[0][0] = 1
    for b in range(1, n + 1):
        nxt = np.zeros(nq, np.int64)
        for ql in range(nq):
            for qs in range(nq):
                nxt[(ql + qs) % nq] += min(left[-1][ql], 10 ** 9) * ds[qs]
        left.append(np.minimum(nxt, 10 ** 9))
    right = [None] * (n + 1)
    right[n] = np.zeros(nq, np.int64)
    right[n][qtot % nq] = 1
    for b in range(n - 1, -1, -1):
        cur = np.zeros(nq, np.int64)
        for ql in range(nq):
            for qs in range(nq):
                cur[ql] += ds[qs] * min(right[b + 1][(ql + qs) % nq], 10 ** 9)
        right[b] = np.minimum(cur, 10 ** 9)
    dims = []
    for b in range(n + 1):
        cap = np.minimum(left[b], right[b])
        if cap.sum() > D:
            base = D // nq
            x = np.minimum(cap, base)
            rem = D - x.sum()
            for q in range(nq):          # hand the remainder to sectors that still have room
                add = min(rem, cap[q] - x[q])
                x[q] += add
                rem -= add
            cap = x
        dims.append(cap.astype(np.int64))
    # a second right-to-left pass keeps every left sector no larger than what its right neighbours can carry
    for b in range(n - 1, -1, -1):
        for ql in range(nq):
            room = sum(ds[qs] * dims[b + 1][(ql + qs) % nq] for qs in range(nq))
            dims[b][ql] = min(dims[b][ql], room)
    return dims


def rho_block_sizes(ss, dims):
    """Orders of the density-matrix blocks of every bond update of one sweep (right half-sweep: rho over (l_{b-1}, s_b); left: over
    (s_{b+1}, l_{b+1})), one block per charge of the middle link -- what the CPU baseline's eigensolver model needs."""
    nq = ss.nq
    sq = np.asarray(ss.charges)
    n = len(dims) - 1
    out = []
    for b in range(1, n):                                  # bond (b, b+1), sites 1-based
        left = [sum(int(dims[b - 1][ql]) for ql in range(nq) for s in sq if (ql + int(s)) % nq == q) for q in range(nq)]
        right = [sum(int(dims[b + 1][qr]) for qr in range(nq) for s in sq if (qr - int(s)) % nq == q) for q in range(nq)]
        out += [x for x in left if x > 0] + [x for x in right if x > 0]
    return out


def random_right_canonical(ss, n, D, cplx, seed):
    nq, d = ss.nq, ss.d
    sq = np.asarray(ss.charges)
    dims = link_profile(ss, n, D)
    labels = [np.concatenate([np.full(int(m), q, np.int32) for q, m in enumerate(dm)]) for dm in dims]
    offs = [np.concatenate([[0], np.cumsum(dm)]) for dm in dims]
    sites = []
    for j in range(n):
        rng = np.random.default_rng(1000 * 7 + j + seed)
        Dl, Dr = int(dims[j].sum()), int(dims[j + 1].sum())
        A = np.zeros((Dl, d, Dr), dtype=complex if cplx else float, order="F")
        for ql in range(nq):
            rows = int(dims[j][ql])
            if rows == 0:
                continue
            cols = [(s, (ql + int(sq[s])) % nq) for s in range(d)]
            width = sum(int(dims[j + 1][qr]) for _, qr in cols)
            M = rng.standard_normal((width, rows)) + (1j * rng.standard_normal((width, rows)) if cplx else 0)
            Q, _ = np.linalg.qr(M)                       # width x rows, orthonormal columns
            B = Q.conj().T                               # orthonormal rows: sum_{s,r} B B^H = 1
            c0 = 0
            for s, qr in cols:
                w = int(dims[j + 1][qr])
                A[offs[j][ql]:offs[j][ql] + rows, s, offs[j + 1][qr]:offs[j + 1][qr] + w] = B[:, c0:c0 + w]
                c0 += w
        sites.append(A)
    return cb.MPS(sites, labels, nq=nq, center=1), dims


def run_sweeps(ctx, H, psi, D, nsweep):
    """nsweep calls of cb200_dmrg with ONE sweep each (what the reference's rep loop does, src/dmrg.h:589-596), maxdim = D binding."""
    out = []
    for sw in range(nsweep):                             # sweep 0 also pays the first-touch growth of the memory pool
        info = cb.DmrgInfo()
        l0 = ctx.launches
        t0 = time.perf_counter()
        en, psi = cb.dmrg(H, [], psi, cb.Sweeps(1, D, float(os.environ.get("CB200_SWEEP_CUTOFF", "1e-12")), niter=2), ctx=ctx, info=info)
        wall = time.perf_counter() - t0
        t = info.timers
        out.append(dict(sweep=sw, wall_s=wall, energy=en, bond_updates=len(info.stats), max_m=max(s["m"] for s in info.stats),
                        env_s=t["env"], matvec_s=t["matvec"], level1_s=t["level1"], trunc_s=t["trunc"], other_s=t["other"],
                        matvecs=int(t["nmatvec"]), matvec_tflops=t["matvec_flops"] / max(t["matvec"], 1e-12) / 1e12, launches=ctx.launches - l0))
    return out


def main():
    st, bc, n, D, cp = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), [float(x) for x in sys.argv[5].split(",")]
    nsweep = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    ss = models.make_siteset(st, n)
    mpo = ss.hamiltonian(bc, n, 2.0, cp)
    H = cb.MPO(mpo.sites, mpo.link_charges, ss.charges, nq=ss.nq)
    t0 = time.perf_counter()
    psi, dims = random_right_canonical(ss, n, D, H.is_complex(), seed=0)
    t_gen = time.perf_counter() - t0
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    ctx = cb.Context(local)
    if world > 1:        # torchrun: every rank drives the same sweep; matvecs of bonds with D >= 1600 are row-sharded, Z_N truncations spread
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        big = max(int(a.sum()) for a in dims)
        ctx.shard(rank, world, mode="p2p", landing_bytes=big * big * ss.d * ss.d * (16 if H.is_complex() else 8) // max(ss.nq, 1) + (1 << 20),
                  min_dim=int(os.environ.get("CB200_SHARD_MIN_DIM", "1600")))
    if os.environ.get("CB200_SWEEP_PACKED", "0") == "1":  # the MPS crosses the boundary in the QDense block layout (cb200_mps.layout = 1)
        psi = cb.PackedMPS.from_dense(psi, ss.charges)
    out = run_sweeps(ctx, H, psi, D, nsweep)
    if rank != 0:
        return
    print(json.dumps(dict(n_gpus=world, workload="%s %s L=%d D=%d j=%s (synthetic right-canonical MPS, true MPO k=%d)" % (st, bc, n, D, sys.argv[5], max(len(q) for q in mpo.link_charges)),
                          dtype="c128" if H.is_complex() else "f64", bond_profile=[int(x.sum()) for x in dims], host_generation_s=t_gen, sweeps=out)))


if __name__ == "__main__":
    main()
