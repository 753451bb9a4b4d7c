"""Physical runs through the stand-in driver on the GPU: seconds per sweep and the phase split the engine reports."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chain_b200 as cb
from chain_b200 import driver

cases = {
    "cfg1": ("golden", "p", 6, 100, 1e-8, 1e-9, "-1", 2.0, 0, 1),
    "cfg2": ("haagerupq", "p", 6, 1600, 1e-8, 1e-2, "0,-1,0", 2.0, 0, 1),
    "cfg3": ("golden", "p", 24, 800, 1e-8, 1e-6, "-1", 2.0, 0, 2),
    "cfg4": ("haagerup", "o", 12, 3000, 1e-8, 1e-4, "0,-1,0", 2.0, 0, 1),
}
name = sys.argv[1]
ctx = cb.Context(0)
sim = driver.Simulation(*cases[name], out="/tmp/drv_" + name, seed=1, ctx=ctx, log=lambda *a: None, checkpoint=False)
t0 = time.perf_counter()
ens, states = sim.run()
wall = time.perf_counter() - t0
nsw = sum(s["nsweep"] for s in sim.sweep_log)
tm = {}
for s in sim.sweep_log:
    for k, v in s["timers"].items():
        tm[k] = tm.get(k, 0.0) + v
# the CPU oracle (NumPy restatement of ITensor's dmrg, all host cores) on the same model and schedule, for a bounded number of sweeps:
# warm-up ladder + reps until ~CPU_BUDGET_S seconds are spent; seconds per sweep of the LAST rep (the bond dimension has grown by then)
cpu = None
if os.environ.get("CB200_DRIVER_CPU", "1") != "0":
    import numpy as np
    from oracle import models as omodels, mpo as ompo, dmrg as od, cpu_baseline as cbz
    st, bc, n, D, cutoff, prec, cpl, U, q, nst = cases[name]
    threads = cbz.use_all_host_threads()
    site = omodels.make_site(st)
    W, qk = ompo.build_mpo(omodels.hamiltonian_terms(st, bc, n, U, driver.parse_couplings(cpl)), site, n)
    c = float(np.float32(cutoff))
    psi = od.product_mps(site, n, q, np.random.default_rng(1), dtype=complex if np.iscomplexobj(W[1]) else float)
    budget = float(os.environ.get("CB200_DRIVER_CPU_BUDGET_S", "40"))
    t0 = time.perf_counter()
    en_c, psi = od.dmrg(W, [], psi, od.Sweeps(5, [10, 20, 40, 80, 200], [max(x, c) for x in (1e-5, 1e-6, 1e-7, 1e-8, 1e-9)]), weight=1000.0)
    warm = time.perf_counter() - t0
    reps, last, bond_dim = 0, warm / 5, 200
    while time.perf_counter() - t0 < budget and reps < 40:
        if bond_dim <= D - 200:
            bond_dim += 200
        t1 = time.perf_counter()
        en_c, psi = od.dmrg(W, [], psi, od.Sweeps(1, bond_dim, c), weight=1000.0)
        last = time.perf_counter() - t1
        reps += 1
    cpu = {"kind": "port (oracle/dmrg.py, NumPy + OpenBLAS; the reference's ITensor C++ build cannot be compiled here)", "threads": threads,
           "warmup_5_sweeps_s": warm, "reps_timed": reps, "s_per_sweep_last_rep": last, "energy": en_c,
           "max_bond": max(a.shape[2] for a in psi.A[:-1])}
print(json.dumps({"case": name, "cpu_oracle": cpu, "energies": ens, "sweeps": nsw, "wall_s": wall, "dmrg_s": sum(s["seconds"] for s in sim.sweep_log),
                  "s_per_sweep": sum(s["seconds"] for s in sim.sweep_log) / nsw, "max_bond": max(max(p.bond_dims()) for p in states),
                  "last_sweep_s": sim.sweep_log[-1]["seconds"], "timers": tm}))
