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
print(json.dumps({"case": name, "energies": ens, "sweeps": nsw, "wall_s": wall, "dmrg_s": sum(s["seconds"] for s in sim.sweep_log),
                  "s_per_sweep": sum(s["seconds"] for s in sim.sweep_log) / nsw, "max_bond": max(max(p.bond_dims()) for p in states),
                  "last_sweep_s": sim.sweep_log[-1]["seconds"], "timers": tm}))
