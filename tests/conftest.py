import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu() -> bool:
    try:
        from kbbq_b200 import api

        return api.lib().kbbq_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle port and the emulation tool are test infrastructure: build them on demand."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "libkbbq_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    emu = os.path.join(ROOT, "tests", "emu", "libkbbq_emu.so")
    src = os.path.join(ROOT, "tests", "emu", "kbbq_emu.cc")
    if not os.path.exists(emu) or os.path.getmtime(emu) < os.path.getmtime(src):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-o", emu, src])
    yield
