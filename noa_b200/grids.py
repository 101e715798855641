"""RNG-free synthetic (kinetic energy, recoil energy) grids of BASELINE.json's configurations
(SURVEY.md section 8(d)).  Built with numpy on the host so the oracle and the GPU see the very
same doubles; the benchmark then moves them to the device (or pins them) once.

  set B ("in-range"): K = 10^(2+4u) GeV, nu = 10^(-2 + (log10(0.5)+2) v), q = nu K  -- every pair
                      is inside the kinematic range of all four processes (throughput/roofline)
  set A ("broad"):    K = 10^(-3+9u), nu = 10^(-6+6v)  -- exercises every early exit and branch
  notebook grid:      K = linspace(1e-3, 1e6, 10000), q = 0.0505 K  (docs/pms/muon_dcs_calc.ipynb:174)
  table energies:     K_i = 10^(-2 + 8 i/(n-1)) GeV
with u_i = (i+0.5)/n and v_i = frac((i+0.5) * golden ratio conjugate).
"""
import numpy as np

_PHI = 0.6180339887498949


def _uv(n, start=0, count=None):
    count = n - start if count is None else count
    i = np.arange(start, start + count, dtype=np.float64)
    u = (i + 0.5) / n
    v = np.modf((i + 0.5) * _PHI)[0]
    return u, v


def set_b(n, start=0, count=None):
    """Slice [start, start+count) of the n-pair in-range grid."""
    u, v = _uv(n, start, count)
    K = 10.0 ** (2.0 + 4.0 * u)
    nu = 10.0 ** (-2.0 + (np.log10(0.5) + 2.0) * v)
    return K, nu * K


def set_a(n, start=0, count=None):
    u, v = _uv(n, start, count)
    K = 10.0 ** (-3.0 + 9.0 * u)
    nu = 10.0 ** (-6.0 + 6.0 * v)
    return K, nu * K


def notebook_grid(n=10000):
    K = np.linspace(1e-3, 1e6, n)
    return K, 0.0505 * K


def table_energies(n=10000, lo=-2.0, hi=6.0):
    i = np.arange(n, dtype=np.float64)
    return 10.0 ** (lo + (hi - lo) * i / max(n - 1, 1))
