"""Three H_eff products on a bench workload -- the process `ncu -k regex:mix_smem` is pointed at (no torch import, starts in seconds)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chain_b200 as cb
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.build_workload(name, seed=2000)
dt = np.complex128 if w["dims"]["cplx"] else np.float64
phi = np.asfortranarray(w["phi"].astype(dt))
ctx = cb.Context(device=0)
bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
               site_charges=w["sq"], nq=w["nq"], ctx=ctx)
for _ in range(3):
    y = bond.apply(phi)
print("done", float(np.abs(y).max()))
