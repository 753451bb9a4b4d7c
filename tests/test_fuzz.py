"""Robustness sweep of svdBond on designed Schmidt spectra (tools/fuzz_truncate.py: exact multiplets, pairs split at rounding level, flat and
rank-deficient spectra, maxdim landing inside a multiplet, ragged / empty Z3 sectors) on the host-emulation build against the oracle.  The
long form (80+ seeds) and the eigensolver sweep (tools/fuzz_heig.py) are run by hand; this keeps a short one in the suite."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_truncation_on_designed_spectra_matches_oracle(emu_ctx, monkeypatch, capsys):
    import fuzz_truncate
    monkeypatch.setattr(sys, "argv", ["fuzz_truncate.py", "25"])
    rc = fuzz_truncate.main()
    out = capsys.readouterr().out
    assert rc == 0 and "failures: 0" in out, out


def test_abi_misuse_returns_status_codes():
    """tools/fuzz_abi.py --quick: every 5th of the ~750 single-field corruptions of a valid cb200_dmrg / entropies / overlap / mpo_element /
    translate call (NULL arrays, contradictory ranks / dtypes / dimensions, labels out of range, maxdim < 1, NaN cutoff ...) comes back with a
    status code and leaves the context usable.  The full list under ASan + UBSan is the tool's default mode."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_abi.py"), "--quick"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "failures: 0" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_dmrg_differential_fuzz_short(emu_ctx, monkeypatch, capsys):
    """tools/fuzz_dmrg.py, 80 draws: whole dmrg() calls on random models / boundary conditions / schedules (niter 1-4, shrinking maxdim, mindim,
    cutoffs down to the true-SVD branch, penalised second states) against the oracle, bond by bond, with the oracle's own sensitivity to a 1e-15
    perturbation of the start as the yardstick."""
    import fuzz_dmrg
    monkeypatch.setattr(sys, "argv", ["fuzz_dmrg.py", "80", "0"])
    rc = fuzz_dmrg.main()
    out = capsys.readouterr().out
    assert rc == 0 and "failures: 0" in out, out


def test_context_reuse_across_models_rebuilds_omega():
    """Regression (found by tools/fuzz_dmrg.py, draws 41 -> 77 in one context): the Omega cache of a context was keyed by a 64-bit multiply-xor
    hash of the two MPO tensors only; the 0/1-valued boundary tensors of two different sine-square-deformed chains collided and the second run
    contracted its last bond with the first run's Omega (energy off by 0.36, no error).  A hit is now confirmed element for element: a run
    in a context that has seen another model is bit-identical to the same run in a fresh context."""
    import numpy as np
    import chain_b200 as cb
    import common
    import fuzz_dmrg
    from conftest import EMU_LIB

    def run(ctx, seed):
        p = fuzz_dmrg.draw(np.random.default_rng(1000 + seed))
        site, W, qk = common.build_model(p["st"], p["bc"], p["n"], p["cp"], U=p["U"])
        H = cb.MPO(W, qk, site.charges, nq=site.nq)
        rng = np.random.default_rng(seed)
        psi0 = common.warm_start(site, W, p["n"], p["q"], seed=seed, maxdim=int(rng.choice([2, 4, 8])))
        info = cb.DmrgInfo()
        e, _ = cb.dmrg(H, [], common.to_cb(psi0), cb.Sweeps(p["nsweep"], p["maxd"], p["cut"], niter=p["niter"], mindim=p["mind"]), weight=p["weight"],
                       ctx=ctx, info=info)
        return e, [(s["m"], s["energy"]) for s in info.stats]

    fresh = run(cb.Context(lib_path=EMU_LIB), 77)
    used = cb.Context(lib_path=EMU_LIB)
    run(used, 41)
    assert run(used, 77) == fresh


def test_measurement_entry_points_on_random_states(emu_ctx, monkeypatch, capsys):
    """tools/fuzz_measure.py, 150 draws: entropies / overlap / exact translation / <a|O|b> of random flux-conserving MPS (unsorted link labels,
    empty sectors, rank-deficient links, both storage layouts) against dense state-vector arithmetic."""
    import fuzz_measure
    monkeypatch.setattr(sys, "argv", ["fuzz_measure.py", "150", "0"])
    rc = fuzz_measure.main()
    out = capsys.readouterr().out
    assert rc == 0 and "failures: 0" in out, out


def test_bond_update_on_random_structures_with_penalties_and_virtual_ranks(emu_ctx, monkeypatch, capsys):
    """tools/fuzz_bond.py, 25 draws: one bond update on random ragged structures incl. complex penalty vectors and 2-8 virtual ranks."""
    import fuzz_bond
    monkeypatch.setattr(sys, "argv", ["fuzz_bond.py", "25", "0"])
    rc = fuzz_bond.main()
    out = capsys.readouterr().out
    assert rc == 0 and "failures: 0" in out, out


def test_sharded_dmrg_is_bit_identical_to_unsharded_three_processes():
    """tools/fuzz_shard.py, 12 draws on 3 real processes (gloo rendezvous, p2p exchange over shared memory, every bond sharded): the sharded
    engine reproduces the unsharded run bit for bit and all ranks stay in lockstep."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_shard.py"), "12", "3", "0"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "failures: 0" in p.stdout and "worst energy gap to the unsharded run 0.00e+00" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
