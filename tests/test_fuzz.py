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
