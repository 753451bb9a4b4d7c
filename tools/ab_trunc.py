"""A/B of one complete bond update between two builds of the library:  python tools/ab_trunc.py WORKLOAD lib1.so lib2.so ..."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chain_b200 as cb
import bench

w = bench.build_workload(sys.argv[1], seed=2000)
dt = np.complex128 if w["dims"]["cplx"] else np.float64
phi = np.asfortranarray(w["phi"].astype(dt))
for lib in sys.argv[2:]:
    ctx = cb.Context(device=0, lib_path=lib)
    bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
                   site_charges=w["sq"], nq=w["nq"], ctx=ctx)
    for rep in range(3):
        bu = bond.update_timed(phi, "left", maxdim=w["dims"]["D"], cutoff=1e-12)
        print(os.path.basename(lib), rep, {k: round(v, 4) if isinstance(v, float) else v for k, v in bu.items()}, flush=True)
    bond.close()
    del ctx
