import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

EMU_LIB = os.path.join(ROOT, "tests", "hostemu", "libcb200_hostemu.so")
PRODUCT_LIB = os.path.join(ROOT, "chain_b200", "libchain_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def emu_ctx():
    """Engine linked against the serial CPU backend: HOST-LOGIC tests only (tests/hostemu, never shipped)."""
    if not os.path.exists(EMU_LIB):
        subprocess.check_call(["make", "-s", "tests/hostemu/libcb200_hostemu.so"], cwd=ROOT)
    import chain_b200 as cb
    return cb.Context(lib_path=EMU_LIB)


@pytest.fixture(scope="session")
def gpu_ctx():
    """The product library on cuda:0. No fallback: a missing library or device is an error, not a skip."""
    import chain_b200 as cb
    return cb.Context(device=0, lib_path=PRODUCT_LIB)
