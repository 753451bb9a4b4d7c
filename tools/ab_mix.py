"""A/B of the Omega-mix kernel between two builds of the library on several bond shapes:

    python tools/ab_mix.py OLD.so NEW.so [WORKLOAD ...]

For every workload the H_eff product of both builds is compared element by element (the two mix kernels accumulate in the same order,
so the vectors must be bit-identical) and the three stage times (L-stage GEMM, mix, R-stage GEMM) are printed with the mix kernel's
achieved HBM rate.  One JSON line per (workload, library) goes to stdout as soon as it is known."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chain_b200 as cb
import bench

# extra shapes: real Z3 with two 32-row CTAs per SM (cfg5 at a third of the size)
bench.WORKLOADS["cfg5s"] = ("haagerupq", "p", 8, [-1.0, 0.0, 0.0], 1200, "haagerupq p j=(-1,0,0) synthetic D=1200 Z3 f64")

old, new = sys.argv[1], sys.argv[2]
names = sys.argv[3:] or ["cfg2", "cfg4s", "cfg3"]
t_start = time.time()
TWO_STAGE_AB = {"cfg2": 3000, "cfg3": 1024}      # block sizes 3200 (complex) / 1600 (real): below the default thresholds 4096 / 2048
for name in names:
    w = bench.build_workload(name, seed=2000)
    dt = np.complex128 if w["dims"]["cplx"] else np.float64
    phi = np.asfortranarray(w["phi"].astype(dt))
    ys = {}
    for tag, lib in (("old", old), ("new", new)):
        ctx = cb.Context(device=0, lib_path=lib)
        bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
                       site_charges=w["sq"], nq=w["nq"], ctx=ctx)
        ys[tag] = bond.apply(phi).copy()
        st_ms, st_fl = bond.stage_times(phi, reps=5)
        by = bond.stage_bytes()
        rec = {"workload": name, "lib": tag, "stage_ms": [round(x, 4) for x in st_ms], "mix_GBps": round(by[1] / (st_ms[1] * 1e-3) / 1e9, 1) if st_ms[1] > 0 else None,
               "mix_frac_of_6551": round(by[1] / (st_ms[1] * 1e-3) / 1e9 / 6551.4, 3) if st_ms[1] > 0 else None,
               "matvec_tflops": round(bond.matvec_flops / (sum(st_ms) * 1e-3) / 1e12, 2) if sum(st_ms) > 0 else None, "t": round(time.time() - t_start, 1)}
        if tag == "new":
            d = np.abs(ys["new"] - ys["old"]).max()
            rec["max_abs_diff_vs_old"] = float(d)
            rec["bit_identical"] = bool(np.array_equal(ys["new"], ys["old"]))
            rec["ref_norm_max"] = float(np.abs(ys["old"]).max())
        print(json.dumps(rec), flush=True)
        if tag == "new" and name in TWO_STAGE_AB:
            # truncation A/B on the same bond: default eigensolver thresholds against the two-stage reduction forced for this block size
            # (CB200_TWOSTAGE_MIN_N is read per call)
            for label, thr in (("default", None), ("two_stage_from_%d" % TWO_STAGE_AB[name], TWO_STAGE_AB[name])):
                if thr is None:
                    os.environ.pop("CB200_TWOSTAGE_MIN_N", None)
                else:
                    os.environ["CB200_TWOSTAGE_MIN_N"] = str(thr)
                for rep in range(3):
                    bu = bond.update_timed(phi, "left", maxdim=w["dims"]["D"], cutoff=1e-12)
                print(json.dumps({"workload": name, "truncation_ab": label, "trunc_s": round(bu["trunc"], 4), "matvec_s": round(bu["matvec"], 4),
                                  "env_s": round(bu["env"], 4), "m": bu["m"], "t": round(time.time() - t_start, 1)}), flush=True)
            os.environ.pop("CB200_TWOSTAGE_MIN_N", None)
        bond.close()
        ctx.close()
        del bond, ctx
    del w, phi, ys
