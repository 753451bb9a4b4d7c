"""The product's host engine (engine.cpp / engine_io.cpp / capi.cpp -- the same sources libchain_b200.so links) under AddressSanitizer +
UBSan + _GLIBCXX_ASSERTIONS, driven through the C ABI by the bodies of the host-logic tests on the emulation backend: bond operations, the
Jacobi fallback of the truncation, DMRG trajectories, one- / two-stage / batched eigensolver drivers, the true-SVD branch, excited states, the
shard planner with virtual ranks, random ragged Z3 structures, edge cases, the packed MPS boundary, translation and rho-defect measurements.
Builds the instrumented library under /tmp (nothing in the tree changes) and re-executes itself with the sanitizer runtime preloaded.

    python tools/asan_engine.py          # prints the sanitizer reports, if any, and ALL OK
"""
import inspect
import os
import pathlib
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = "/tmp/cb200_asan/libemu_asan.so"


def build():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    src = [os.path.join(ROOT, "chain_b200", "csrc", f) for f in ("engine.cpp", "engine_io.cpp", "capi.cpp")] + [os.path.join(ROOT, "tests", "hostemu", "backend_emu.cpp")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                           "-D_GLIBCXX_ASSERTIONS", "-shared", "-o", LIB] + src)


class Env:   # the two monkeypatch calls the test bodies use
    def setenv(self, k, v):
        os.environ[k] = v

    def delenv(self, k, raising=False):
        os.environ.pop(k, None)


def run():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import chain_b200 as cb
    import common
    import test_edge_cases as te
    import test_random_structures as trs
    import test_rho_defect as tr
    import test_shard as ts
    import test_translation as tt
    ctx, mp = cb.Context(lib_path=LIB), Env()
    for case in common.BOND_CASES:
        common.check_bond_ops(ctx, case)
        common.check_bond_blocks(ctx, case)
    common.check_truncation_survives_ns_failure(ctx, common.BOND_CASES[4], mp)
    for case in common.DMRG_CASES[:2]:
        common.check_dmrg_trajectory(ctx, case, nsweep=8)
    for cplx in (False, True):
        common.check_heig_two_stage(ctx, cplx, 300, 120, mp)
        mp.delenv("CB200_TWOSTAGE_MIN_N")
        common.check_heig_batched(ctx, cplx, [(70, 30), (130, 130), (33, 5)])
        common.check_true_svd_branch(ctx, cplx, "left", D=32)
    common.check_excited_qn(ctx, nsweep=12)
    for case in (common.BOND_CASES[2], common.BOND_CASES[4]):
        for world in (2, 3, 8):
            ts.test_virtual_ranks_sum_to_full_matvec(ctx, case, world)
    ts.test_small_bonds_stay_on_one_gpu(ctx)
    ts.test_eigensolver_refuses_exchange_free_sharding(ctx)
    for seed in range(8):
        try:
            trs.test_random_z3_structures(ctx, seed)
        except BaseException as e:   # noqa: BLE001  (pytest.skip of an empty draw)
            if "Skipped" not in type(e).__name__:
                raise
    te.test_ragged_and_empty_sectors(ctx)
    te.test_bond_dimension_one(ctx)
    te.test_invalid_inputs_give_status_codes(ctx)
    for case in (common.DMRG_CASES[0], common.DMRG_CASES[3], common.DMRG_CASES[4]):
        common.check_packed_mps_boundary(ctx, case)
        common.check_packed_mps_boundary(ctx, case, shuffle=True)
        common.check_mps_boundary_paths(ctx, case, mp)
    for mod in (tt, tr):
        for name, fn in inspect.getmembers(mod, inspect.isfunction):
            if not name.startswith("test_") or any(m.name == "parametrize" for m in getattr(fn, "pytestmark", [])):
                continue
            kw = {"emu_ctx": ctx, "tmp_path": pathlib.Path(tempfile.mkdtemp()), "monkeypatch": mp}
            fn(**{p: kw[p] for p in inspect.signature(fn).parameters})
    print("ALL OK")


if __name__ == "__main__":
    if os.environ.get("CB200_ASAN_CHILD") == "1":
        run()
    else:
        build()
        gcc = lambda f: subprocess.check_output(["gcc", "-print-file-name=" + f], text=True).strip()
        env = dict(os.environ, CB200_ASAN_CHILD="1", LD_PRELOAD=gcc("libasan.so") + " " + gcc("libstdc++.so.6"),
                   ASAN_OPTIONS="detect_leaks=0:halt_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))
