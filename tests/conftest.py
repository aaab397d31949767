import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The product library, built in-tree by nvcc (cross-compiles without a GPU)."""
    from covo_mpc_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def host_check_lib():
    """g++ build of the product's __host__ __device__ model headers (test infrastructure only)."""
    src = os.path.join(ROOT, "tests", "host_check", "model_check.cpp")
    out = os.path.join(ROOT, "tests", "host_check", "libmodel_check.so")
    deps = [src] + [os.path.join(ROOT, "covo_mpc_b200", "csrc", f) for f in ("quad_model.cuh", "hessian_local.cuh", "pid.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, src])
    import ctypes

    return ctypes.CDLL(out)
