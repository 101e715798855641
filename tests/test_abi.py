"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares,
refuses to run without a device, and the Python mirror validates its arguments."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols(header="noa_dcs_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(noa_dcs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from noa_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/noa_dcs_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in noa_b200/_lib.py"
    assert lib.noa_dcs_abi_version() == 3
    assert b"invalid" in lib.noa_dcs_strerror(-1)
    # measurement hooks are not part of the product library any more
    for name in ("noa_dcs_set_pair_mode", "noa_dcs_set_exchange_fence_mode",
                 "noa_dcs_set_max_blocks_per_sm", "noa_dcs_fp64_probe"):
        assert not hasattr(lib, name), f"{name} must live in libnoa_dcs_b200_probe.so"
    probe = _lib.load_probe()
    for name in _declared_symbols("noa_dcs_b200_probe.h"):
        assert hasattr(probe, name), f"{name} declared in noa_dcs_b200_probe.h but not exported"
        assert name in _lib.PROBE_SIGNATURES, name


def test_host_libm_selfcheck():
    """noa_dcs_selfcheck: the host's exp / log / log10 / pow are the ones the kernels restate (it
    also runs at load time; this asserts the count explicitly)."""
    import ctypes
    from noa_b200 import _lib
    lib = _lib.load()
    bad = ctypes.c_int64(-1)
    assert lib.noa_dcs_selfcheck(ctypes.byref(bad)) == 0
    assert bad.value == 0
    assert b"libm" in lib.noa_dcs_strerror(-5)


def test_table_workspace_size_is_pure_arithmetic():
    """noa_dcs_table_workspace_doubles needs no device: row parameters (2 x double2 per row), the
    queue words, and 16 B per node and process with the node count rounded up to whole 6-point
    cells."""
    from noa_b200 import _lib
    lib = _lib.load()
    assert lib.noa_dcs_table_workspace_doubles(0, 180) == 0
    assert lib.noa_dcs_table_workspace_doubles(10, 0) == 0
    for n, mp, nodes in ((1, 1, 6), (10000, 1000, 1002), (1250, 180, 180), (37, 4000, 4002)):
        assert lib.noa_dcs_table_workspace_doubles(n, mp) == 4 * n + 2 + 8 * n * nodes


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from noa_b200 import _lib, dcs, STANDARD_ROCK, MUON_MASS
    with pytest.raises(_lib.NoaDcsError):
        _lib.require_device()
    K = torch.ones(4, dtype=torch.float64)
    with pytest.raises((ValueError, _lib.NoaDcsError)):
        dcs.map(dcs.bremsstrahlung)(K, K, STANDARD_ROCK, MUON_MASS)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "noa_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hh", ".cc", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "dcs_oracle" not in text and "libnoa_ref" not in text, f


def test_api_tokens_and_validation():
    import torch
    from noa_b200 import dcs
    assert [p.index for p in dcs.PROCESSES] == [0, 1, 2, 3]
    with pytest.raises(TypeError):
        dcs.vmap(lambda *a: 0.0)
    with pytest.raises(TypeError):
        dcs.vmap_integral(dcs.bremsstrahlung)
    ri = dcs.recoil_integral(dcs.pair_production, dcs.cel_integrand)
    assert ri.process.index == 1 and ri.integrand.index == 1
    assert dcs.pvmap is dcs.vmap and dcs.pmap is dcs.map
    bad = torch.ones(4, dtype=torch.float32)
    with pytest.raises((ValueError, Exception)):
        dcs.cuda.vmap_bremsstrahlung(bad, bad, bad, (22., 0.1364e-6, 11), 0.10565839)


def test_libtorch_boundary_has_no_python_dependency_and_cli_refuses_cpu():
    """libnoa_dcs_b200_torch.so is what a C++ user links: it must not need libpython.  The
    benchmark executable built on it (benchmark/measure_dcs_calc_cuda.cc) must fail loudly without
    a GPU instead of falling back to anything."""
    import subprocess
    import torch
    so = os.path.join(ROOT, "noa_b200", "libnoa_dcs_b200_torch.so")
    exe = os.path.join(ROOT, "noa_b200", "measure_dcs_calc_cuda")
    if not (os.path.exists(so) and os.path.exists(exe)):
        pytest.skip("LibTorch boundary not built here (python noa_b200/csrc/build_torch_ext.py)")
    syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert " U Py" not in syms and "_Py_" not in syms
    for name in ("vmap_bremsstrahlung", "map_bremsstrahlung", "vmap_integral", "tables",
                 "coulomb_data", "coulomb_transport", "hard_scattering", "soft_scattering"):
        assert name in syms, name
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU path" in r.stderr
