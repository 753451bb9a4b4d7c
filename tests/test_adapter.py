"""The reference-side binding include/chain_b200_itensor.h (drop-in for itensor::dmrg, /root/reference/src/dmrg.h:544,563,596):
type-checked against tests/itensor_stub and EXECUTED -- stub ITensor objects -> adapter -> C ABI -> engine -> stub MPS -- on the
host-emulation library here and on the product library on the GPU box.  Energies must equal the Python mirror's on the same input."""
import os
import subprocess

import numpy as np
import pytest

import chain_b200 as cb
from chain_b200 import models
from conftest import EMU_LIB, PRODUCT_LIB, ROOT

STUB = os.path.join(ROOT, "tests", "itensor_stub")
HDR = os.path.join(ROOT, "include", "chain_b200_itensor.h")


def test_adapter_type_checks_against_the_itensor_stub():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter", "-I" + STUB, "-I" + os.path.join(ROOT, "include"),
                        "-x", "c++", "-include", HDR, os.devnull], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_macro_only_touches_the_three_call_sites():
    """`#define dmrg(...)` is function-like: in the reference sources the token `dmrg` followed by `(` occurs only at the three calls."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference sources not on this machine")
    import re
    hits = []
    for fn in sorted(os.listdir(ref)):
        for i, line in enumerate(open(os.path.join(ref, fn), errors="replace"), 1):
            code = line.split("//")[0]
            if re.search(r"(?<![A-Za-z0-9_])dmrg\s*\(", code):
                hits.append((fn, i))
    assert hits == [("dmrg.h", 544), ("dmrg.h", 563), ("dmrg.h", 596)]


def _build_runner(lib, tag):
    exe = os.path.join(ROOT, "build", "adapter_run_" + tag)
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(STUB, "adapter_run.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(HDR), os.path.getmtime(os.path.join(STUB, "itensor", "all.h"))):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + STUB, "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                               lib, "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def _run_case(ctx, lib, tag, tmp_path):
    n = 6
    ss = models.make_siteset("golden", n)
    mpo = ss.hamiltonian("p", n, 2.0, [-1.0])
    rng = np.random.default_rng(7)
    vecs = [rng.standard_normal(ss.d) for _ in range(n)]
    maxdim, cutoff = [10, 20, 40, 80, 200], [1e-5, 1e-6, 1e-7, 1e-8, 1e-9]
    path = os.path.join(tmp_path, "case.txt")
    with open(path, "w") as f:
        f.write("%d %d %d\n" % (n, ss.d, len(maxdim)))
        for m, c in zip(maxdim, cutoff):
            f.write("%d %.17g\n" % (m, c))
        f.write(" ".join(str(W.shape[0]) for W in mpo.sites) + " 1\n")
        for W in mpo.sites:
            f.write(" ".join("%.17g" % v for v in np.asarray(W, dtype=float).reshape(-1, order="F")) + "\n")
        for v in vecs:
            f.write(" ".join("%.17g" % x for x in v) + "\n")
    out = subprocess.run([_build_runner(lib, tag), path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    en = [float(l.split()[1]) for l in out.stdout.splitlines() if l.startswith("energy")]
    nrm = [float(l.split()[1]) for l in out.stdout.splitlines() if l.startswith("norm1")]
    # the same two calls through the Python mirror
    H = cb.MPO(mpo.sites, mpo.link_charges, ss.charges, nq=1)
    psi0 = cb.MPS([v.reshape(1, ss.d, 1) for v in vecs], [np.zeros(1, np.int32)] * (n + 1), nq=1, center=0)
    sw = cb.Sweeps(len(maxdim), maxdim, cutoff)
    e0, p0 = cb.dmrg(H, [], psi0, sw, weight=1000.0, ctx=ctx)
    e1, _ = cb.dmrg(H, [p0], psi0, sw, weight=1000.0, ctx=ctx)
    assert len(en) == 2 and abs(en[0] - e0) <= 1e-12 * abs(e0) and abs(en[1] - e1) <= 1e-12 * abs(e1), (en, e0, e1)
    assert abs(en[0] - (-4.698543184702)) < 1e-4              # SURVEY App. C, golden p L=6 (five warm-up sweeps only)
    assert all(abs(x - 1.0) < 1e-12 for x in nrm)              # orthogonality centre at site 1, normalised


def test_adapter_runs_end_to_end_on_hostemu(emu_ctx, tmp_path):
    _run_case(emu_ctx, EMU_LIB, "emu", str(tmp_path))


@pytest.mark.gpu
def test_adapter_runs_end_to_end_on_the_gpu(gpu_ctx, tmp_path):
    _run_case(gpu_ctx, PRODUCT_LIB, "gpu", str(tmp_path))
