"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys, and the
GPU arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "heff_matvec_fp64_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_gpu_arm_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_bench_library_call_sequence_on_host_emulation(emu_ctx):
    """The calls bench.py's GPU arm makes, with the arguments it makes them with (workload builder, Bond, dense and block-packed / slice-wise
    products, one timed bond update at cutoff 1e-12 with maxdim = D, per-stage timing, one whole synthetic sweep through cb200_dmrg), at
    smoke sizes on the host emulation: the bench can only run on a GPU box, so a change of the C ABI's argument checks that broke it would
    otherwise first be seen at round end."""
    import numpy as np
    import bench
    import chain_b200 as cb
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_bench as sb
    from chain_b200 import models
    saved = dict(bench.WORKLOADS)
    try:
        bench.WORKLOADS["tiny"] = ("haagerupq", "p", 8, [0.0, -1.0, 0.0], 24, "smoke Z3 c128")
        bench.WORKLOADS["tinyd"] = ("haagerup", "o", 12, [0.0, -1.0, 0.0], 12, "smoke dense f64")
        for name, Ls in (("tiny", 8), ("tinyd", None)):
            w = bench.build_workload(name, seed=2000)
            bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
                           site_charges=w["sq"], nq=w["nq"], ctx=emu_ctx)
            dt = np.complex128 if w["dims"]["cplx"] else np.float64
            phi = np.asfortranarray(w["phi"].astype(dt))
            y = bond.apply(phi)
            b0, cnt = bond.phi_blocks_slice()
            xb = bond.pack(phi)[b0:b0 + cnt].copy()
            yb = np.empty_like(xb)
            bond.apply_blocks_sharded(xb, out=yb)
            assert np.array_equal(bond.unpack(yb), y)
            assert bond.matvec_flops > 0 and all(b > 0 for b in bond.stage_bytes())
            bu = bond.update_timed(phi, "left", maxdim=w["dims"]["D"], cutoff=1e-12)
            assert bu["m"] == w["dims"]["D"] and np.isfinite(bu["energy"])
            bond.stage_times(phi, reps=1)
            bond.close()
            if Ls:
                st, bc, _, cp, D, _ = bench.WORKLOADS[name]
                ss = models.make_siteset(st, Ls)
                mpo = ss.hamiltonian(bc, Ls, 2.0, cp)
                Hs = cb.MPO(mpo.sites, mpo.link_charges, ss.charges, nq=ss.nq)
                psi_s, _ = sb.random_right_canonical(ss, Ls, D, Hs.is_complex(), seed=0)
                sw = sb.run_sweeps(emu_ctx, Hs, psi_s, D, 2)[-1]
                assert sw["bond_updates"] == 2 * (Ls - 1) and sw["max_m"] == D and np.isfinite(sw["energy"])
    finally:
        bench.WORKLOADS.clear()
        bench.WORKLOADS.update(saved)
