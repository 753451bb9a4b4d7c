"""Variants of the Omega-mix kernel inside ONE process (the knobs are read per call):  python tools/ab_mix_variants.py [WORKLOAD ...]
Prints one JSON line per (workload, variant): stage times, achieved HBM rate of the mix, bit-identity against the round-1 kernel."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chain_b200 as cb
import bench

VARIANTS = [("v1", {"CB200_MIX": "v1"}), ("v2", {}), ("v2_cluster4", {"CB200_MIX_CLUSTER": "4"}), ("v2_cluster2", {"CB200_MIX_CLUSTER": "2"}),
            ("v2_rb32", {"CB200_MIX_RB": "32"}), ("v2_t256", {"CB200_MIX_THREADS": "256"}), ("v2_cluster8", {"CB200_MIX_CLUSTER": "8"})]
KNOBS = ["CB200_MIX", "CB200_MIX_CLUSTER", "CB200_MIX_RB", "CB200_MIX_THREADS"]
names = sys.argv[1:] or ["cfg2", "cfg4s"]
t0 = time.time()
ctx = cb.Context(device=0)
for name in names:
    w = bench.build_workload(name, seed=2000)
    dt = np.complex128 if w["dims"]["cplx"] else np.float64
    phi = np.asfortranarray(w["phi"].astype(dt))
    ref = None
    for tag, env in VARIANTS:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
                       site_charges=w["sq"], nq=w["nq"], ctx=ctx)
        y = bond.apply(phi)
        if ref is None:
            ref = y.copy()
        st_ms, _ = bond.stage_times(phi, reps=5)
        by = bond.stage_bytes()
        print(json.dumps({"workload": name, "variant": tag, "mix_ms": round(st_ms[1], 4), "mix_GBps": round(by[1] / (st_ms[1] * 1e-3) / 1e9, 1),
                          "frac_of_6551": round(by[1] / (st_ms[1] * 1e-3) / 1e9 / 6551.4, 3), "gemm_ms": [round(st_ms[0], 3), round(st_ms[2], 3)],
                          "bit_identical_to_v1": bool(np.array_equal(y, ref)), "t": round(time.time() - t0, 1)}), flush=True)
        bond.close()
        del bond, y
    for k in KNOBS:
        os.environ.pop(k, None)
    del w, phi, ref
