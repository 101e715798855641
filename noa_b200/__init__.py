"""noa_b200 -- B200-native muon DCS hot path of NOA/PMS behind the reference's own API.

Layout (only what the path needs):
  csrc/                  CUDA kernels + C ABI (libnoa_dcs_b200.so), LibTorch C++ boundary
  _lib.py                ctypes binding of the C ABI (include/noa_dcs_b200.h)
  physics.py             noa::pms types and constants (AtomicElement, STANDARD_ROCK, MUON_MASS...)
  dcs.py                 noa::pms::dcs mirror: vmap/map/pvmap/pmap, recoil_integral,
                         vmap_integral, dcs.cuda.vmap_* / map_*, fused and table builders
  muons.py               the notebook extension surface (docs/pms/muon_dcs.{cc,cu})
  sharding.py            one-process-per-GPU partitioning + table all-gather

There is no CPU implementation in this package: every entry point needs the CUDA library and a
B200 and raises otherwise.
"""
from . import physics  # noqa: F401
from .physics import (AtomicElement, Material, STANDARD_ROCK, MUON_MASS, ELECTRON_MASS,  # noqa: F401
                      TAU_MASS, X_FRACTION, HYDROGEN, OXYGEN, IRON, LEAD, WATER, ROCK)
from . import dcs  # noqa: F401

__version__ = "0.1.0"
