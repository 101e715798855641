import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- shared fixtures ------------------------------------------------------------------------------
MUON_MASS = 0.10565839
ELEMENTS = {
    "rock": (22., 0.1364E-6, 11),
    "H": (1.0087, 19.2E-9, 1),
    "O": (15.999, 95.0E-9, 8),
    "Fe": (55.845, 286E-9, 26),
    "Pb": (207.2, 823E-9, 82),
}
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.load_port()


@pytest.fixture(scope="session")
def reference():
    """The compiled reference (oracle/_ref) or None where it was never built."""
    import oracle
    return oracle.load_reference()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(GOLDEN_DIR, "dcs_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def special():
    """Zeros, subnormals, thresholds, q >= K, huge values, inf, NaN, negative energies -- outputs of
    the compiled reference (tests/golden/make_special_golden.py)."""
    return np.load(os.path.join(GOLDEN_DIR, "special_golden.npz"))


SPECIAL_ELEMENTS = ("rock", "H", "Pb")


def wild_inputs(n, seed=20261017):
    """(K, q) with no physics in mind: half raw 64-bit patterns reinterpreted as doubles (NaNs,
    infinities, subnormals, both signs), half log-uniform magnitudes over the whole double range
    with a random sign.  Seeded; the reference's answer on them is whatever its arithmetic gives."""
    rng = np.random.default_rng(seed)

    def draw():
        raw = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64).view(np.float64)
        with np.errstate(all="ignore"):
            mag = 10.0 ** rng.uniform(-320, 308, size=n) * rng.choice([-1.0, 1.0], size=n,
                                                                      p=[0.1, 0.9])
        return np.where(rng.random(n) < 0.5, raw, mag)

    return draw(), draw()


@pytest.fixture(scope="session")
def hostcheck():
    """Host build of the kernels' scalar arithmetic (test fixture, see oracle/hostcheck.cc)."""
    out = os.path.join(ROOT, "oracle", "_build", "libhostcheck.so")
    src = os.path.join(ROOT, "oracle", "hostcheck.cc")
    deps = [src] + [os.path.join(ROOT, "noa_b200", "csrc", f)
                    for f in ("dcs_math.cuh", "folded_ops.cuh", "coulomb_math.cuh", "dcs_params.hh", "glibm.cuh",
                              "glibm_tables.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-fPIC", "-shared",
                        src, "-o", out], check=True)
    lib = ctypes.CDLL(out)
    for name in ("hostcheck_coulomb_data", "hostcheck_coulomb_transport",
                 "hostcheck_hard_scattering", "hostcheck_soft_scattering"):
        getattr(lib, name).restype = ctypes.c_int
    lib.hostcheck_glibm.restype = ctypes.c_int64
    lib.hostcheck_glibm.argtypes = [ctypes.c_int64, ctypes.c_uint64]
    return lib
