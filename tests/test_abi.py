"""CPU checks of the C-ABI boundary: the product library loads, exports every symbol include/chain_b200.h declares,
and refuses to run without a device (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from common import ROOT, PRODUCT_LIB


def _built():
    if not os.path.exists(PRODUCT_LIB):
        subprocess.check_call(["make", "-s", "chain_b200/libchain_b200.so"], cwd=ROOT)


def test_header_symbols_all_exported():
    _built()
    import chain_b200._lib as L
    hdr = open(os.path.join(ROOT, "include", "chain_b200.h")).read()
    declared = set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", hdr))
    bound = {name for name, _, _ in L.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    lib = L.load(PRODUCT_LIB)       # raises if any symbol is missing
    for name in declared:
        assert hasattr(lib, name)


def test_no_oracle_or_cpu_path_in_product():
    # the product package must not import the oracle, and the shared library must not contain the host-emulation backend
    for fn in os.listdir(os.path.join(ROOT, "chain_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "chain_b200", fn)).read()
            assert "oracle" not in src.replace("the oracle", "")
    _built()
    out = subprocess.run(["nm", "-D", "--defined-only", PRODUCT_LIB], capture_output=True, text=True).stdout
    assert "cb200_create" in out
    strings = subprocess.run(["strings", PRODUCT_LIB], capture_output=True, text=True).stdout
    assert "host-emulation" not in strings
    assert "cuda-sm_100a" in strings


def test_create_fails_loudly_without_gpu():
    _built()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this checks the GPU-less behaviour")
    import chain_b200 as cb
    with pytest.raises(cb.ChainB200Error) as e:
        cb.Context(device=0, lib_path=PRODUCT_LIB)
    assert "no CPU fallback" in str(e.value)


def test_missing_library_is_an_error():
    import chain_b200 as cb
    with pytest.raises(cb.ChainB200Error):
        cb.Context(lib_path="/nonexistent/libchain_b200.so")
