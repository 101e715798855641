"""CPU pre-flight for the GPU parity tests: the kernels' scalar arithmetic, compiled for the host
(oracle/hostcheck.cc), must agree BIT FOR BIT with (1) the system libm for exp/log/log10 and
(2) the oracle for the four DCS and the closed-form ionisation integrals."""
import ctypes

import numpy as np

from conftest import ELEMENTS, MUON_MASS
from noa_b200 import grids

_dp = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def test_glibm_is_bit_exact_against_system_libm(hostcheck):
    assert hostcheck.hostcheck_glibm(3_000_000, 20251017) == 0


def test_kernel_math_on_host_matches_oracle_bit_for_bit(hostcheck, port):
    for kind, (K, q) in (("A", grids.set_a(1 << 14)), ("B", grids.set_b(1 << 14))):
        for en, el in ELEMENTS.items():
            for p in range(4):
                out = np.zeros_like(K)
                rc = hostcheck.hostcheck_dcs(p, _p(K), _p(q), _p(out), ctypes.c_int64(K.size),
                                             ctypes.c_double(el[0]), ctypes.c_double(el[1]),
                                             ctypes.c_int32(el[2]), ctypes.c_double(MUON_MASS))
                assert rc == 0
                want = port.vmap(p, K, q, el, MUON_MASS, threads=4)
                assert np.array_equal(out, want, equal_nan=True), (kind, en, p)


def test_pair_production_with_hoisted_row_part_on_host(hostcheck, port, special):
    """The table kernels compute pair production's Lorentz factor and zeta once per row and hand
    them to every node's evaluation (dcs_math.cuh: PairRow): the same bits as the oracle, on the
    synthetic sets and on the special-value grid (thresholds, zeros, huge values, NaN)."""
    from conftest import SPECIAL_ELEMENTS
    cases = [(grids.set_a(1 << 14), ELEMENTS, None), (grids.set_b(1 << 14), ELEMENTS, None),
             ((special["S_K"], special["S_q"]), {e: ELEMENTS[e] for e in SPECIAL_ELEMENTS}, special)]
    for (K, q), elements, golden in cases:
        for en, el in elements.items():
            out = np.zeros_like(K)
            rc = hostcheck.hostcheck_pair_with_row_part(
                _p(K), _p(q), _p(out), ctypes.c_int64(K.size), ctypes.c_double(el[0]),
                ctypes.c_double(el[1]), ctypes.c_int32(el[2]), ctypes.c_double(MUON_MASS))
            assert rc == 0
            want = port.vmap(1, K, q, el, MUON_MASS, threads=4) if golden is None \
                else golden[f"vmap_S_{en}_pair_production"]
            assert np.array_equal(out, want, equal_nan=True), en


def test_closed_form_ionisation_on_host(hostcheck, port):
    K = grids.table_energies(256, -2.0, 1.0)     # below the 10.8 GeV switch (dcs.hh:964)
    el = ELEMENTS["rock"]
    for ig in (0, 1):
        out = np.zeros_like(K)
        hostcheck.hostcheck_ionisation_closed_form(ig, _p(K), _p(out), ctypes.c_int64(K.size),
                                                   ctypes.c_double(0.05), ctypes.c_double(el[0]),
                                                   ctypes.c_double(el[1]), ctypes.c_int32(el[2]),
                                                   ctypes.c_double(MUON_MASS))
        want = port.vmap_integral(3, ig, K, 0.05, 180, el, MUON_MASS)
        assert np.array_equal(out, want)


def test_kernel_math_on_host_matches_reference_on_special_values(hostcheck, special):
    from conftest import SPECIAL_ELEMENTS
    K, q = special["S_K"], special["S_q"]
    for en in SPECIAL_ELEMENTS:
        el = ELEMENTS[en]
        for p, pn in enumerate(("bremsstrahlung", "pair_production", "photonuclear", "ionisation")):
            out = np.zeros_like(K)
            rc = hostcheck.hostcheck_dcs(p, _p(K), _p(q), _p(out), ctypes.c_int64(K.size),
                                         ctypes.c_double(el[0]), ctypes.c_double(el[1]),
                                         ctypes.c_int32(el[2]), ctypes.c_double(MUON_MASS))
            assert rc == 0
            want = special[f"vmap_S_{en}_{pn}"]
            bad = ~((out == want) | (np.isnan(out) & np.isnan(want)))
            assert not bad.any(), (en, pn, K[bad][:4], q[bad][:4], out[bad][:4], want[bad][:4])


def test_kernel_math_on_host_with_tau_projectile(hostcheck, port):
    """Every mass-dependent hoisted invariant (dcs_params.hh) with a projectile other than the
    muon: tau, physics.hh:59."""
    tau = 1.77682
    for kind, (K, q) in (("A", grids.set_a(1 << 13)), ("B", grids.set_b(1 << 13))):
        for en in ("rock", "H", "Pb"):
            el = ELEMENTS[en]
            for p in range(4):
                out = np.zeros_like(K)
                hostcheck.hostcheck_dcs(p, _p(K), _p(q), _p(out), ctypes.c_int64(K.size),
                                        ctypes.c_double(el[0]), ctypes.c_double(el[1]),
                                        ctypes.c_int32(el[2]), ctypes.c_double(tau))
                want = port.vmap(p, K, q, el, tau, threads=4)
                assert np.array_equal(out, want, equal_nan=True), (kind, en, p)


def test_kernel_math_on_host_on_wild_inputs(hostcheck, port):
    from conftest import wild_inputs
    K, q = wild_inputs(1 << 16)
    for en in ("rock", "H"):
        el = ELEMENTS[en]
        for p in range(4):
            out = np.zeros_like(K)
            hostcheck.hostcheck_dcs(p, _p(K), _p(q), _p(out), ctypes.c_int64(K.size),
                                    ctypes.c_double(el[0]), ctypes.c_double(el[1]),
                                    ctypes.c_int32(el[2]), ctypes.c_double(MUON_MASS))
            with np.errstate(all="ignore"):
                want = port.vmap(p, K, q, el, MUON_MASS, threads=4)
            assert ((out == want) | (np.isnan(out) & np.isnan(want))).all(), (en, p)
