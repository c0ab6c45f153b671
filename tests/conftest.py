"""Shared fixtures.  `oracle` is the CPU checker (oracle/libsonar_oracle.so); `gpu` is the product
library (sonido-sonar_b200/libsonar.so) and exists only in tests marked `gpu`."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_LIB = os.path.join(ROOT, "oracle", "libsonar_oracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("sonido-sonar_b200")


@pytest.fixture(scope="session")
def capi(pkg):
    return pkg.capi


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def oracle(capi):
    if not os.path.exists(ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = capi.SonarLib(ORACLE_LIB)
    assert lib.backend == "cpu-oracle"
    return lib


@pytest.fixture(scope="session")
def gpu(capi):
    lib = capi.SonarLib()  # raises if libsonar.so is missing or no CUDA device: no CPU fallback
    assert lib.backend == "cuda-sm100a"
    yield lib
    lib.close()
