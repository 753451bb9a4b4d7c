#!/usr/bin/env python
"""bench.py -- H_eff matvec throughput of the Chain DMRG hot path on B200 (BASELINE.json metric).

A "step" is one application of the two-site effective Hamiltonian (ITensor LocalMPO::product, the call made three
times per bond by davidson at /root/reference/src/dmrg.h:544,563,596) on the synthetic form of BASELINE config 2:
haagerupq periodic, j=(0,-1,0), the model's true compressed MPO (k = 26 = Z3 sectors 10/8/8) at a central bond,
bond dimension D = 1600 as Z3 sectors (534,533,533), complex128, block-sparse.  Environments and phi are seeded N(0,1)
in the allowed charge blocks.

  value  : algorithmic FP64 TFLOP/s (F_alg of SURVEY 8d: block-sparse, sparse-W flops only) with phi resident in HBM
  e2e    : the same through the C-ABI call cb200_bond_apply_blocks with pinned HOST buffers in ITensor's QDense block storage
           (H2D phi + D2H result per step); e2e.dense_layout = cb200_bond_apply on dense arrays with zeros
  roofline: the DMMA GEMM kernel (both D^3 stages) against the FP64 peak measured live (cuBLAS DGEMM 8192^3)
  bond_update / sweep: one complete two-site bond update resident on the device, and ONE whole DMRG sweep through cb200_dmrg (host MPS in/out)
  cpu_baseline / --impl reference: the oracle's port of ITensor's block-pair permute+zgemm loop on the host cores

N > 1 (torchrun): ONE matvec row-sharded over the N GPUs (strong scaling): every rank holds phi, applies
H_eff for its slice of the bra rows, and the output rows are all-gathered by peer stores fused into the R-stage GEMM
epilogue (--exchange p2p, default) or summed with ncclAllReduce (--exchange nccl).  value = F_alg of the whole matvec
divided by the slowest rank's time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_emit = None

WORKLOADS = {
    # name: (site, bc, L, couplings, D, description)
    "cfg2": ("haagerupq", "p", 8, [0.0, -1.0, 0.0], 1600, "haagerupq p j=(0,-1,0) synthetic D=1600 Z3(534,533,533) k=26(10,8,8) c128 (BASELINE configs[1])"),
    "cfg5": ("haagerupq", "p", 8, [-1.0, 0.0, 0.0], 4000, "haagerupq p j=(-1,0,0) synthetic D=4000 Z3(1334,1333,1333) k=26(10,8,8) f64 (BASELINE configs[4])"),
    "cfg3": ("golden", "p", 24, [-1.0], 800, "golden p synthetic D=800 k=10 f64 dense (BASELINE configs[2])"),
    "cfg4": ("haagerup", "o", 12, [0.0, -1.0, 0.0], 3000, "haagerup o synthetic D=3000 k=14 f64 dense (BASELINE configs[3])"),
    "cfg4s": ("haagerup", "o", 12, [0.0, -1.0, 0.0], 1500, "haagerup o synthetic D=1500 k=14 f64 dense (half-size BASELINE configs[3])"),
    "tiny": ("haagerupq", "p", 8, [0.0, -1.0, 0.0], 96, "haagerupq p j=(0,-1,0) synthetic D=96 (smoke size)"),
}


def sectors(D, nq):
    base, rem = divmod(D, nq)
    return [base + (1 if q < rem else 0) for q in range(nq)]


def labels(dims):
    return np.concatenate([np.full(n, q, np.int32) for q, n in enumerate(dims)])


def build_workload(name, seed):
    """Synthetic bond problem: true MPO tensors of the model, N(0,1) in the allowed blocks of L, R, phi."""
    from chain_b200 import models
    st, bc, L, cp, D, desc = WORKLOADS[name]
    ss = models.make_siteset(st, L)
    mpo = ss.hamiltonian(bc, L, 2.0, cp)
    b = L // 2                                # central bond: sites b, b+1
    W1, W2 = mpo.sites[b - 1], mpo.sites[b]
    qkl, qkm, qkr = mpo.link_charges[b - 1], mpo.link_charges[b], mpo.link_charges[b + 1]
    nq = ss.nq
    cplx = mpo.is_complex()
    dl = sectors(D, nq)
    ql = labels(dl)
    qr = (ql + 0) % nq                        # right link: same sector sizes
    sq = ss.charges
    rng = np.random.default_rng(seed)

    def rnd(shape, mask):
        x = rng.standard_normal(shape)
        if cplx:
            x = x + 1j * rng.standard_normal(shape)
        return np.asfortranarray(x * mask)
    Lm = (ql[:, None, None] - qkl[None, :, None] - ql[None, None, :]) % nq == 0          # q(l') = q(l) + q(a)
    Rm = (qr[:, None, None] - qkr[None, :, None] - qr[None, None, :]) % nq == 0
    Lenv, Renv = rnd((D, len(qkl), D), Lm), rnd((D, len(qkr), D), Rm)
    pm = (ql[:, None, None, None] + sq[None, :, None, None] + sq[None, None, :, None] - qr[None, None, None, :]) % nq == 0
    phi = rnd((D, ss.d, ss.d, D), pm)
    phi /= np.linalg.norm(phi)
    dims = dict(Dl=dl, Dr=dl, kl=np.bincount(qkl, minlength=nq).tolist(), km=np.bincount(qkm, minlength=nq).tolist(),
                kr=np.bincount(qkr, minlength=nq).tolist(), ds=np.bincount(sq, minlength=nq).tolist(), d=ss.d, nq=nq, cplx=bool(cplx), D=D)
    return dict(L=Lenv, W1=W1, W2=W2, R=Renv, phi=phi, ql=ql, qr=qr, qkl=qkl, qkm=qkm, qkr=qkr, sq=sq, nq=nq, desc=desc, dims=dims)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_ev = gpu, [], threading.Event()

    def run(self):
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def summary(self):
        self.stop_ev.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


def run_reference(args, rank):
    """--impl reference: the CPU path (oracle port of ITensor's block-pair loop), one bounded sample per step."""
    if rank != 0:
        return
    from oracle import cpu_baseline as cbz
    st, bc, L, cp, D, desc = WORKLOADS[args.workload]
    w = build_dims_only(args.workload)
    budget = max(2.0, min(20.0, 60.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        cbz.matvec_sample(w["Dl"], w["Dr"], w["kl"], w["kr"], w["ds"], w["cplx"], budget_s=budget / 4)
    fl, secs, sample = 0.0, 0.0, ""
    for s in range(args.steps):
        r = cbz.matvec_sample(w["Dl"], w["Dr"], w["kl"], w["kr"], w["ds"], w["cplx"], budget_s=budget, seed=s)
        fl += r["flops"]; secs += r["seconds"]; sample = r["sample"]
    v = fl / secs / 1e12
    line = {"impl": "reference", "metric": "heff_matvec_fp64_tflops", "value": v, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128" if w["cplx"] else "f64", "data": "synthetic",
            "config": {"workload": desc, "note": "each step = a bounded sample of the matvec's block-pair GEMMs on the host cores"},
            "cpu_baseline": {"value": v, "unit": "TFLOP/s", "cores": cbz.host_threads(), "host_cores": cbz.host_cores(), "kind": "port", "sample": sample + " per step"},
            "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def build_dims_only(name):
    from chain_b200 import models
    st, bc, L, cp, D, desc = WORKLOADS[name]
    ss = models.make_siteset(st, L)
    mpo = ss.hamiltonian(bc, L, 2.0, cp)
    b = L // 2
    nq = ss.nq
    return dict(Dl=sectors(D, nq), Dr=sectors(D, nq), kl=np.bincount(mpo.link_charges[b - 1], minlength=nq).tolist(),
                kr=np.bincount(mpo.link_charges[b + 1], minlength=nq).tolist(), ds=np.bincount(ss.charges, minlength=nq).tolist(),
                cplx=bool(mpo.is_complex()))


def main():
    # stdout carries exactly ONE JSON line: everything libraries print there (e.g. "NCCL version ...") is sent to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    global _emit
    _emit = emit
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the whole-sweep measurement (the `sweep` key)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: how the ranks exchange output rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import chain_b200 as cb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = build_workload(args.workload, seed=2000)          # every rank holds the same bond problem
    ctx = cb.Context(device=local_rank)
    exchange = args.exchange
    if world > 1:
        try:
            ctx.shard(rank, world, mode=exchange, landing_bytes=w["phi"].size * (16 if w["dims"]["cplx"] else 8), min_dim=0)
        except cb.ChainB200Error as err:
            if exchange != "p2p":
                raise
            print("bench: %s -- falling back to --exchange nccl" % err, file=sys.stderr)   # same decision on every rank (see Context.shard)
            exchange = "nccl"
            ctx.shard(rank, world, mode=exchange, min_dim=0)
    bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
                   site_charges=w["sq"], nq=w["nq"], ctx=ctx)
    falg = bond.matvec_flops
    dt = np.complex128 if w["dims"]["cplx"] else np.float64
    phi = np.asfortranarray(w["phi"].astype(dt))

    # FP64 peaks measured live (MEASURED_PEAKS.json carries no FP64 number)
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, a)
    torch.cuda.synchronize()
    cublas = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, a); e1.record(); torch.cuda.synchronize()
        cublas = max(cublas, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a
    torch.cuda.empty_cache()
    dmma_pipe = cb.k_fp64_peak(0, 20000, ctx=ctx)

    # ---- timed region 1: device-resident matvecs (CUDA events on the library's launching stream)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    l0 = ctx.launches
    t0 = time.perf_counter()
    ms = bond.apply_timed(phi, reps=args.steps, warmup=args.warmup)
    barrier()
    wall = time.perf_counter() - t0
    launches = (ctx.launches - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.summary()
    ms_t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = falg / (ms_max * 1e-3) / 1e12                 # falg is the WHOLE matvec; the ranks split it

    # ---- per-stage times -> roofline of the dominant kernel (the DMMA GEMM: L-stage + R-stage launches)
    st_ms, st_fl = bond.stage_times(phi, reps=3)
    gemm_ms = st_ms[0] + st_ms[2]
    gemm_fl = st_fl[0] + st_fl[2]
    roofline = {"kernel": "gemm_grouped_kernel (L-stage + R-stage launches of one matvec)", "bound": "tensor", "achieved": gemm_fl / (gemm_ms * 1e-3) / 1e12,
                "peak": cublas, "unit": "TFLOP/s", "frac": gemm_fl / (gemm_ms * 1e-3) / 1e12 / cublas, "traffic": None,
                "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "dmma_issue_loop_tflops": dmma_pipe, "share_of_step": gemm_ms / sum(st_ms),
                "stage_ms": {"gemm_L": st_ms[0], "mix": st_ms[1], "gemm_R": st_ms[2]},
                "stage_flops": {"gemm_L": st_fl[0], "mix": st_fl[1], "gemm_R": st_fl[2]}}

    if w["dims"]["cplx"] and os.environ.get("CB200_CPLX_3M", "96") != "0":
        # complex GEMMs run the 3-multiplication product: `achieved` counts ALGORITHMIC flops (4 real multiplies per complex MAC, the
        # contract's definition), the DMMA pipe executes 3 -- so achieved can exceed a DGEMM-rate peak; pipe_frac is what the pipe sees
        roofline["executed_over_algorithmic_flops"] = 0.75
        roofline["pipe_frac"] = roofline["frac"] * 0.75
        roofline["note"] = "c128 stages use the 3M complex product (3 DMMA per complex MAC): frac is algorithmic flops / cuBLAS DGEMM rate, pipe_frac = 0.75 * frac is the executed share of that rate"

    # ---- timed region 2: end to end through the C ABI with pinned host buffers (H2D + compute + D2H every step)
    xh = torch.empty(phi.size * (2 if w["dims"]["cplx"] else 1), dtype=torch.float64, pin_memory=True)
    yh = torch.empty_like(xh, pin_memory=True)
    xn = xh.numpy().view(dt).reshape(phi.shape, order="F")
    yn = yh.numpy().view(dt).reshape(phi.shape, order="F")
    xn[...] = phi
    for _ in range(2):
        bond.apply(xn, out=yn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bond.apply(xn, out=yn)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_dense_v = falg * args.steps / float(e2e_t.item()) / 1e12
    e2e_dense_ms = e2e_s / args.steps * 1e3
    nbytes_dense = int(phi.size * phi.itemsize)
    del xh, yh, xn, yn

    # the headline e2e: the same call on ITensor's own QN storage (QDense block list, cb200_bond_apply_blocks) -- only the
    # charge-allowed blocks cross PCIe; for a dense model (nq = 1) the two layouts coincide
    # N > 1: rank g carries only ITS slice of the block-packed vector over PCIe (cb200_bond_apply_blocks_sharded): the slices are
    # all-gathered between the GPUs over NVLink, and the result crosses PCIe once (1/N per rank) instead of N times
    nb = bond.phi_blocks_elems
    b0, cnt = bond.phi_blocks_slice()
    xbh = torch.empty(max(cnt, 1) * (2 if w["dims"]["cplx"] else 1), dtype=torch.float64, pin_memory=True)
    ybh = torch.empty_like(xbh, pin_memory=True)
    xb, yb = xbh.numpy().view(dt)[:cnt], ybh.numpy().view(dt)[:cnt]
    xb[...] = bond.pack(phi)[b0:b0 + cnt]
    for _ in range(2):
        bond.apply_blocks_sharded(xb, out=yb)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bond.apply_blocks_sharded(xb, out=yb)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_v = falg * args.steps / float(e2e_t.item()) / 1e12
    nbytes = int(nb * phi.itemsize)

    # ---- one complete bond update resident on the device (davidson -> svdBond -> makeL): the unit a DMRG sweep is made of
    for _ in range(2):      # two warm-ups: the first grows the memory pool, the second still re-shapes the allocator's block cache
        bond.update_timed(phi, "left", maxdim=w["dims"]["D"], cutoff=1e-12)    # (A/B in round 2: call 2 of a process 0.45-0.56 s, calls 1, 3, 4.. 0.29 s)
    barrier()
    bu = bond.update_timed(phi, "left", maxdim=w["dims"]["D"], cutoff=1e-12)
    barrier()
    bu_tot = bu["matvec"] + bu["level1"] + bu["trunc"] + bu["env"]
    bu_t = torch.tensor([bu_tot], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(bu_t, op=dist.ReduceOp.MAX)
    L_chain = WORKLOADS[args.workload][2]
    bond_update = {"matvec_s": bu["matvec"], "davidson_level1_s": bu["level1"], "truncation_s": bu["trunc"], "env_update_s": bu["env"],
                   "total_s": float(bu_t.item()), "kept_states": bu["m"], "matvecs": 3,
                   "sweep_s_if_all_bonds_at_D": 2 * (L_chain - 1) * float(bu_t.item()),
                   "note": "one two-site bond update (3 matvecs, Gram-Schmidt, truncation to maxdim=D, environment update) with everything "
                           "resident in HBM; the sweep figure is 2(L-1) such updates, an upper bound since the edge bonds are smaller"}

    # ---- the other half of the BASELINE metric: DMRG sweep wall time.  ONE whole sweep (2(L-1) bond updates) through cb200_dmrg --
    # host MPS in, host MPS out -- on a seeded random right-canonical MPS with the capped bond profile of an L = 12 chain of the same
    # model at the same D (tools/sweep_bench.py); sweep 0 warms the memory pool, sweep 1 is reported.  N = 1 only (the scaling of whole
    # sweeps is in profiles/, the context here is sharded with min_dim = 0 for the matvec measurement)
    sweep = None
    if world == 1 and not args.no_sweep and args.workload in ("cfg2", "cfg3", "tiny"):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sweep_bench as sb
        from chain_b200 import models
        st, bc, _, cp, D, _ = WORKLOADS[args.workload]
        Ls = {"cfg2": 12, "cfg3": 24, "tiny": 8}[args.workload]
        ss = models.make_siteset(st, Ls)
        mpo = ss.hamiltonian(bc, Ls, 2.0, cp)
        Hs = cb.MPO(mpo.sites, mpo.link_charges, ss.charges, nq=ss.nq)
        psi_s, dims_s = sb.random_right_canonical(ss, Ls, D, Hs.is_complex(), seed=0)
        sw = sb.run_sweeps(ctx, Hs, psi_s, D, 2)[-1]
        del psi_s
        sweep = {"wall_s": sw["wall_s"], "workload": "%s %s L=%d D=%d, synthetic right-canonical MPS, true MPO; one sweep = one cb200_dmrg call with host MPS in/out" % (st, bc, Ls, D),
                 "bond_updates": sw["bond_updates"], "bond_profile": [int(x.sum()) for x in dims_s], "matvec_s": sw["matvec_s"], "truncation_s": sw["trunc_s"],
                 "env_update_s": sw["env_s"], "davidson_level1_s": sw["level1_s"], "other_s": sw["other_s"], "matvecs": sw["matvecs"],
                 "matvec_tflops": sw["matvec_tflops"], "gpu_launches": sw["launches"], "energy": sw["energy"],
                 "matvec_flops": sw["matvec_tflops"] * 1e12 * sw["matvec_s"],
                 # environment update = one D^3 stage pair per site instead of per site pair: matvec flops of one product / d
                 "env_flops_est": sw["matvec_tflops"] * 1e12 * sw["matvec_s"] / max(sw["matvecs"], 1) * sw["bond_updates"] / ss.d,
                 "rho_block_sizes": sb.rho_block_sizes(ss, dims_s)}

    # DRAM traffic of the dominant kernel: a constant read from the committed ncu --set full summary of this workload (it cannot be
    # measured outside a profiler); the summary says which launches it covers
    traffic, traffic_src = None, None
    for rnd in ("r02", "r01"):
        try:
            with open(os.path.join(ROOT, "profiles", "%s_ncu_gemm_%s_summary.json" % (rnd, args.workload))) as f:
                js = json.load(f)
            traffic = js.get("dram_bytes_per_matvec_gemm_launches")
            traffic_src = "profiles/%s_ncu_gemm_%s_summary.json: %s" % (rnd, args.workload, js.get("dram_bytes_covers", "dram__bytes_read.sum + dram__bytes_write.sum of the L-stage launch ONLY (the R-stage counters were not captured)"))
            if traffic:
                break
        except Exception:
            pass
    roofline["traffic"] = traffic
    roofline["traffic_source"] = traffic_src
    by = bond.stage_bytes()
    roofline["algorithmic_bytes"] = {"gemm_L": by[0], "mix": by[1], "gemm_R": by[2]}

    # ---- HBM-bound kernels of the path against the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs; north_star asks for both rooflines)
    hbm_peak, hbm_src = 6551.4, "fallback (MEASURED_PEAKS.json absent)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    es = phi.itemsize
    nvec = 3
    l1 = cb.k_level1_timed(bond.phi_blocks_elems, nv=nvec, cplx=bool(w["dims"]["cplx"]), reps=10, ctx=ctx)

    def hbm(kernel, nbytes, ms, what):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None
        return {"kernel": kernel, "bytes": nbytes, "ms": ms, "GB/s": gbs, "frac": gbs / hbm_peak if gbs else None, "what": what}
    roofline_hbm = {"peak_GB/s": hbm_peak, "peak_source": hbm_src, "kernels": [
        hbm("mix_smem2_kernel", by[1], st_ms[1], "Omega = W_b (x) W_{b+1} applied between the two GEMM stages: |T1| read + |T2| written"),
        hbm("dot_partial+dot_final (multi_dot)", (nvec + 1) * bond.phi_blocks_elems * es, l1[0], "Davidson Gram-Schmidt / penalty overlaps: %d dots in one pass over y" % nvec),
        hbm("lincomb_kernel", (nvec + 2) * bond.phi_blocks_elems * es, l1[1], "Davidson Ritz vector / residual / projector update: dst += sum of %d vectors" % nvec),
        hbm("device copy", 2 * bond.phi_blocks_elems * es, l1[2], "Krylov vector copy (cudaMemcpyAsync D2D)")]}

    # DRAM traffic of the mix launch from the committed ncu --set full summary (cannot be measured outside a profiler), like roofline.traffic
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_mix2_%s_summary.json" % args.workload)) as f:
            mj = json.load(f)
        roofline_hbm["kernels"][0]["traffic"] = mj["dram_total_GB"] * 1e9
        roofline_hbm["kernels"][0]["traffic_source"] = "profiles/r02_ncu_mix2_%s_summary.json (N=1 launch: dram__bytes_read.sum + dram__bytes_write.sum; %s)" % (
            args.workload, "stalls: long scoreboard on the entry-list loads, L1 hit %.0f %%" % mj["l1_hit_rate_pct"])
    except Exception:
        pass

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import cpu_baseline as cbz     # the one place bench.py touches the oracle: the reported CPU baseline
            d = w["dims"]
            r = cbz.matvec_sample(d["Dl"], d["Dr"], d["kl"], d["kr"], d["ds"], d["cplx"], budget_s=12.0)
            cpu = {"value": r["tflops"], "unit": "TFLOP/s", "cores": cbz.host_threads(), "host_cores": cbz.host_cores(), "kind": "port", "sample": r["sample"]}
            if sweep is not None:
                # the CPU sweep time beside the GPU sweep: same bond profile, same flops, LAPACK *syev per rho block
                est = cbz.sweep_estimate(sweep["matvec_flops"], sweep["env_flops_est"], sweep["rho_block_sizes"], d["cplx"], r["tflops"])
                cpu["sweep_s"] = est["seconds"]
                cpu["sweep_parts"] = {"contraction_s": est["contraction_s"], "eigensolver_s": est["eigensolver_s"]}
                cpu["sweep_how"] = est["how"]
                sweep["cpu_sweep_s"] = est["seconds"]
                sweep.pop("rho_block_sizes", None)
        line = {"metric": "heff_matvec_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "c128" if w["dims"]["cplx"] else "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "dims": w["dims"], "f_alg_per_matvec": falg,
                           "l2": "inputs larger than L2 (phi + environments + intermediates >> 126 MB)",
                           "parallelism": "single GPU" if world == 1 else "matvec row-sharded x%d, exchange=%s" % (world, exchange)},
                "clocks": clocks,
                "e2e": {"value": e2e_v, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_s / args.steps * 1e3,
                        "call": ("cb200_bond_apply_blocks: pinned host vectors in ITensor's QDense block storage, H2D + product + D2H every step" if world == 1 else
                                 "cb200_bond_apply_blocks_sharded: each rank uploads/downloads its 1/%d slice of the block-packed vector (pinned), slices all-gathered over NVLink; bytes are the whole job's" % world),
                        "dense_layout": {"value": e2e_dense_v, "ms_per_step": e2e_dense_ms, "h2d_bytes_per_step": nbytes_dense, "d2h_bytes_per_step": nbytes_dense,
                                         "call": "cb200_bond_apply: dense [l,s1,s2,r] arrays with zeros in the forbidden blocks"}},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "bond_update": bond_update, "sweep": sweep, "wall_s_timed_region": wall}
        emit(line)
    bond.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
