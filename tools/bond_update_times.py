"""Phase times of one complete bond update (Davidson + truncation + environment update) on a bench workload."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import chain_b200 as cb  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
maxdim = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name][4]
w = bench.build_workload(name, seed=2000)
ctx = cb.Context(0)
bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
               site_charges=w["sq"], nq=w["nq"], ctx=ctx)
r = bond.update_timed(w["phi"], "left", maxdim=maxdim, cutoff=1e-12)
r["workload"] = w["desc"]
print(json.dumps(r))
