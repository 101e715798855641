"""The reference's notebook extension surface (docs/pms/muon_dcs.cu:8-16, docs/pms/muon_dcs.cc:8-45)
on the GPU: `muons.bremsstrahlung(K, q)` etc. for standard rock and muons, float64 tensors in (CUDA,
or CPU as in the reference's own extension -- those run through the host-buffer path), tensor out
on the same device, GIL released.  This module is the compiled pybind11 extension noa_b200/_muons.so, which
calls the C++ LibTorch boundary (noa::pms::dcs::cuda::*, csrc/torch_api.cc) -> C ABI -> kernels.
"""
import os

import torch  # noqa: F401  (libtorch must be loaded before the extension)

_HERE = os.path.dirname(os.path.abspath(__file__))
if not os.path.exists(os.path.join(_HERE, "_muons.so")):
    raise ImportError("noa_b200/_muons.so is missing: run `python noa_b200/csrc/build_torch_ext.py` "
                      "(or __graft_entry__.build())")

from ._muons import (bremsstrahlung, pair_production, photonuclear, ionisation,  # noqa: E402,F401
                     all_processes, tables, recoil_integral, water, serialise,
                     coulomb_hard_scattering, soft_scattering, reference_call_sites)
