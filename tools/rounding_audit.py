"""Rounding audit of the shared parity-test bodies (no GPU needed).

The `-m gpu` tests and the host-logic tests run the SAME bodies (tests/common.py) against the oracle; the only thing that differs between the two
backends, as far as a tolerance is concerned, is rounding: the product library sums GEMM tiles in another order than the serial host emulation, and
its complex kernels use the 3M product (normwise, not componentwise, error).  The host emulation can imitate exactly that
(tests/hostemu/backend_emu.cpp: CB200_EMU_ROUNDING=<seed> perturbs every GEMM result by u * ulps * 2^-53 * sum_k|a_k b_k| / sqrt(K) per real
component; CB200_EMU_ROUNDING_ULPS scales it), so the margin of every tolerance against "a different but equally valid arithmetic" can be measured
here -- in particular for GPU tests that were written after the round's GPU minutes were spent (tests/test_zz_gpu_converged.py).

  python tools/rounding_audit.py converged [NSWEEP] [SEEDS...]     gaps of check_excited_qn_converged per seed (1, 16 ulps)
  python tools/rounding_audit.py suite [ULPS...]                   the host-logic files under pytest, one run per noise scale (default 1 8 64)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SUITE = ["tests/test_host_logic.py", "tests/test_edge_cases.py", "tests/test_random_structures.py", "tests/test_shard.py", "tests/test_translation.py",
         "tests/test_rho_defect.py", "tests/test_driver.py", "tests/test_fuzz.py"]


def converged(nsweep, seeds):
    import common
    import chain_b200 as cb
    from conftest import EMU_LIB
    ctx = cb.Context(lib_path=EMU_LIB)
    worst = {}
    for ulps in ("1", "16"):
        for seed in seeds:
            os.environ["CB200_EMU_ROUNDING"] = str(seed)
            os.environ["CB200_EMU_ROUNDING_ULPS"] = ulps
            m, t, ok = [], time.time(), True
            try:
                common.check_excited_qn_converged(ctx, nsweep=nsweep, measured=m)
            except AssertionError:
                ok = False
            for g in m:
                for k, v in g.items():
                    worst[k] = max(worst.get(k, 0.0), v)
            print(json.dumps({"seed": seed, "ulps": ulps, "ok": ok, "s": round(time.time() - t, 1), "gaps": m}), flush=True)
    tol = {"energy_rel": 1e-10, "ed_abs": 1e-9, "overlap": 1e-9, "ee": 1e-8, "prev_overlap": 1e-6}
    print(json.dumps({"worst": worst, "worst_over_tolerance": {k: worst[k] / tol[k] for k in worst}}))


def suite(ulps_list):
    rc = 0
    for ulps in ulps_list:
        env = dict(os.environ, CB200_EMU_ROUNDING="5", CB200_EMU_ROUNDING_ULPS=str(ulps))
        p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *SUITE, "--deselect",
                            "tests/test_host_logic.py::test_excited_state_in_qn_sector_converged"], cwd=ROOT, env=env, capture_output=True, text=True)
        tail = [l for l in p.stdout.splitlines() if l.startswith("FAILED") or " passed" in l or " failed" in l]
        print("ulps %s: %s" % (ulps, " | ".join(tail)), flush=True)
        rc |= p.returncode if float(ulps) <= 8 else 0
    return rc


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "suite"
    if mode == "converged":
        converged(int(sys.argv[2]) if len(sys.argv) > 2 else 120, [int(x) for x in sys.argv[3:]] or [1, 2, 3])
    else:
        sys.exit(suite(sys.argv[2:] or ["1", "8", "64"]))
