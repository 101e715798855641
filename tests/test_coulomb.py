"""Coulomb scattering + soft scattering (the reference's dcs.hh:499-952).

CPU part (runs everywhere): the C oracle (oracle/coulomb_oracle.c) and the host build of the
kernels' arithmetic (oracle/hostcheck.cc) against tests/golden/coulomb_golden.npz, which was
generated from the compiled reference, and against the compiled reference itself where present.
GPU part (-m gpu): the CUDA kernels through the C ABI / Python mirror against the same fixtures and
against the oracle on larger grids.  Everything is compared BIT FOR BIT (NaN == NaN)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ELEMENTS, GOLDEN_DIR, MUON_MASS
from noa_b200 import grids

MU_TAGS = {"one": np.array([1.0]), "small": np.array([1e-3]), "tiny": np.array([1e-12])}
WATER_W = np.array([0.111894, 0.888106])[:, None]


@pytest.fixture(scope="module")
def cg():
    return np.load(os.path.join(GOLDEN_DIR, "coulomb_golden.npz"))


def same(a, b):
    return np.array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1), equal_nan=True)


def mu_cases(cg):
    return list(MU_TAGS.items()) + [("grid", cg["C_mu_grid"])]


# ---- CPU: oracle port --------------------------------------------------------------------------
def test_port_matches_golden(port, cg):
    K = cg["C_K"]
    for en, el in ELEMENTS.items():
        fcm, scr, fspin, invl = port.coulomb_data(K, el, MUON_MASS)
        assert same(fcm, cg[f"cd_{en}_fcm"]) and same(scr, cg[f"cd_{en}_screen"]), en
        assert same(fspin, cg[f"cd_{en}_fspin"]) and same(invl, cg[f"cd_{en}_invlambda"]), en
        for tag, mu in mu_cases(cg):
            assert same(port.coulomb_transport(scr, fspin, mu), cg[f"ct_{en}_{tag}"]), (en, tag)
        mu0, lbh = port.hard_scattering(cg[f"ct_{en}_one"], fcm, scr, invl, fspin)
        assert same(mu0, cg[f"hs_{en}_mu0"]) and same(lbh, cg[f"hs_{en}_lbh"]), en
        assert same(port.soft_scattering(K, el, MUON_MASS), cg[f"ss_{en}"]), en
    st = lambda name: np.stack((cg[f"cd_H_{name}"], cg[f"cd_O_{name}"]))
    mu0, lbh = port.hard_scattering(np.stack((cg["ct_H_one"], cg["ct_O_one"])), st("fcm"),
                                    st("screen"), cg["hs_water_invlambda"], st("fspin"))
    assert same(mu0, cg["hs_water_mu0"]) and same(lbh, cg["hs_water_lbh"])


def test_port_matches_compiled_reference(port, reference):
    if reference is None or "soft_scattering" not in reference._coulomb:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    K = grids.table_energies(1500, -3.5, 7.0)
    for el in (ELEMENTS["rock"], ELEMENTS["H"], ELEMENTS["Pb"]):
        for mass in (MUON_MASS, 1.77682):
            a, b = port.coulomb_data(K, el, mass), reference.coulomb_data(K, el, mass)
            assert all(same(x, y) for x, y in zip(a, b))
            mu = 10.0 ** np.linspace(-16, 0.2, K.size)
            G = port.coulomb_transport(a[1], a[2], mu)
            assert same(G, reference.coulomb_transport(b[1], b[2], mu))
            G1 = port.coulomb_transport(a[1], a[2], [1.0])
            ha = port.hard_scattering(G1, a[0], a[1], a[3], a[2])
            hb = reference.hard_scattering(G1, b[0], b[1], b[3], b[2])
            assert same(ha[0], hb[0]) and same(ha[1], hb[1])
            assert same(port.soft_scattering(K, el, mass), reference.soft_scattering(K, el, mass))


def test_values_are_physical(cg):
    """Sanity of the fixtures themselves: the reference test's structure
    (test/unit/test-dcs-calc.cc:134-178) -- screening factors positive, the hard-scattering cutoff
    within (0, MAX_MU0], mean free paths positive, soft scattering positive above threshold."""
    max_mu0 = 0.5 * (1. - np.cos(np.pi / 180.))
    for en in ELEMENTS:
        assert np.all(cg[f"cd_{en}_screen"][:, :3] > 0)
        assert np.all(cg[f"cd_{en}_invlambda"] > 0)
        mu0, lbh = cg[f"hs_{en}_mu0"], cg[f"hs_{en}_lbh"]
        assert np.all(mu0 >= 0) and np.all(mu0 <= max_mu0) and np.all(lbh > 0)
        assert np.all(cg[f"ss_{en}"][cg["C_K"] > 1.0] > 0)


# ---- CPU: the kernels' arithmetic compiled for the host ------------------------------------------
_dp = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def test_kernel_math_on_host_matches_golden(hostcheck, cg):
    K = np.ascontiguousarray(cg["C_K"])
    n = K.size
    i64, i32, f64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    for en, (A, I, Z) in ELEMENTS.items():
        fcm, scr = np.zeros((n, 2)), np.zeros((n, 9))
        fspin, invl = np.zeros(n), np.zeros(n)
        hostcheck.hostcheck_coulomb_data(_p(K), i64(n), f64(A), f64(I), i32(Z), f64(MUON_MASS),
                                         _p(fcm), _p(scr), _p(fspin), _p(invl))
        assert same(fcm, cg[f"cd_{en}_fcm"]) and same(scr, cg[f"cd_{en}_screen"]), en
        assert same(fspin, cg[f"cd_{en}_fspin"]) and same(invl, cg[f"cd_{en}_invlambda"]), en
        for tag, mu in mu_cases(cg):
            mu = np.ascontiguousarray(mu, dtype=np.float64)
            coef = np.zeros((n, 2))
            hostcheck.hostcheck_coulomb_transport(_p(scr), _p(fspin), _p(mu), i64(mu.size), i64(n),
                                                  _p(coef))
            assert same(coef, cg[f"ct_{en}_{tag}"]), (en, tag)
        G = np.ascontiguousarray(cg[f"ct_{en}_one"])
        mu0, lbh = np.zeros(n), np.zeros(n)
        hostcheck.hostcheck_hard_scattering(_p(G), _p(fcm), _p(scr), _p(invl), _p(fspin), i32(1),
                                            i64(n), _p(mu0), _p(lbh))
        assert same(mu0, cg[f"hs_{en}_mu0"]) and same(lbh, cg[f"hs_{en}_lbh"]), en
        ms1 = np.zeros(n)
        hostcheck.hostcheck_soft_scattering(_p(K), i64(n), f64(A), f64(I), i32(Z), f64(MUON_MASS),
                                            _p(ms1))
        assert same(ms1, cg[f"ss_{en}"]), en
    st = lambda name: np.ascontiguousarray(np.stack((cg[f"cd_H_{name}"], cg[f"cd_O_{name}"])))
    G = np.ascontiguousarray(np.stack((cg["ct_H_one"], cg["ct_O_one"])))
    invl = np.ascontiguousarray(cg["hs_water_invlambda"])
    mu0, lbh = np.zeros(n), np.zeros(n)
    fcm, scr, fspin = st("fcm"), st("screen"), st("fspin")
    hostcheck.hostcheck_hard_scattering(_p(G), _p(fcm), _p(scr), _p(invl), _p(fspin), i32(2), i64(n),
                                        _p(mu0), _p(lbh))
    assert same(mu0, cg["hs_water_mu0"]) and same(lbh, cg["hs_water_lbh"])


# ---- GPU -----------------------------------------------------------------------------------------
def _gpu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _run_gpu_chain(K, el, mass, mu_list):
    """coulomb_data -> coulomb_transport (each mu) -> hard_scattering (mu = 1) -> soft_scattering,
    with the reference's call shapes (test/unit/test-dcs-calc.cc:134-178)."""
    import torch
    from noa_b200 import dcs
    Kd = _gpu(K)
    n = Kd.numel()
    z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device="cuda")
    fCM, screen, fspin, invlambda = z(n, 2), z(n, 9), z(n), z(n)
    dcs.coulomb_data(fCM, screen, fspin, invlambda, Kd, el, mass)
    Gs = []
    for mu in mu_list:
        G = z(n, 2)
        dcs.coulomb_transport(G, screen, fspin, _gpu(mu))
        Gs.append(G)
    G1 = z(n, 2)
    dcs.coulomb_transport(G1, screen, fspin, torch.tensor(1.0, dtype=torch.float64, device="cuda"))
    mu0, lb_h = z(n), z(n)
    dcs.hard_scattering(mu0, lb_h, G1.view(1, n, 2), fCM.view(1, n, 2), screen.view(1, n, 9),
                        invlambda.view(1, n), fspin.view(1, n))
    ms1 = z(n)
    dcs.soft_scattering(ms1, Kd, el, mass)
    c = lambda t: t.cpu().numpy()
    return (c(fCM), c(screen), c(fspin), c(invlambda)), [c(G) for G in Gs], (c(mu0), c(lb_h)), c(ms1)


@pytest.mark.gpu
def test_gpu_matches_golden(cg):
    K = cg["C_K"]
    for en, el in ELEMENTS.items():
        data, Gs, hs, ms1 = _run_gpu_chain(K, el, MUON_MASS, [mu for _, mu in mu_cases(cg)])
        for got, name in zip(data, ("fcm", "screen", "fspin", "invlambda")):
            assert same(got, cg[f"cd_{en}_{name}"]), (en, name)
        for got, (tag, _) in zip(Gs, mu_cases(cg)):
            assert same(got, cg[f"ct_{en}_{tag}"]), (en, tag)
        assert same(hs[0], cg[f"hs_{en}_mu0"]) and same(hs[1], cg[f"hs_{en}_lbh"]), en
        assert same(ms1, cg[f"ss_{en}"]), en


@pytest.mark.gpu
def test_gpu_two_element_hard_scattering(cg):
    import torch
    from noa_b200 import dcs
    st = lambda name: _gpu(np.stack((cg[f"cd_H_{name}"], cg[f"cd_O_{name}"])))
    n = cg["C_K"].size
    mu0 = torch.zeros(n, dtype=torch.float64, device="cuda")
    lb_h = torch.zeros_like(mu0)
    dcs.hard_scattering(mu0, lb_h, _gpu(np.stack((cg["ct_H_one"], cg["ct_O_one"]))), st("fcm"),
                        st("screen"), _gpu(cg["hs_water_invlambda"]), st("fspin"))
    assert same(mu0.cpu().numpy(), cg["hs_water_mu0"])
    assert same(lb_h.cpu().numpy(), cg["hs_water_lbh"])


@pytest.mark.gpu
def test_gpu_matches_oracle_on_large_grids(port):
    """10^4 energies (BASELINE table grid, extended range), tau projectile included, ragged size."""
    for n, el, mass in ((10007, ELEMENTS["rock"], MUON_MASS), (4099, ELEMENTS["Pb"], 1.77682),
                        (1, ELEMENTS["H"], MUON_MASS)):
        K = grids.table_energies(n, -3.5, 7.0)
        mu = 10.0 ** np.linspace(-16, 0.2, n)
        data, Gs, hs, ms1 = _run_gpu_chain(K, el, mass, [mu])
        want = port.coulomb_data(K, el, mass)
        assert all(same(g, w) for g, w in zip(data, want)), (n, "coulomb_data")
        assert same(Gs[0], port.coulomb_transport(want[1], want[2], mu)), (n, "transport")
        G1 = port.coulomb_transport(want[1], want[2], [1.0])
        wmu0, wlbh = port.hard_scattering(G1, want[0], want[1], want[3], want[2])
        assert same(hs[0], wmu0) and same(hs[1], wlbh), (n, "hard_scattering")
        assert same(ms1, port.soft_scattering(K, el, mass)), (n, "soft_scattering")


@pytest.mark.gpu
def test_cxx_boundary_matches_golden(cg):
    """The same chain through noa::pms::dcs::cuda::* (LibTorch boundary, pybind module)."""
    from noa_b200 import muons
    Kd = _gpu(cg["C_K"])
    fCM, screen, fspin, invlambda, G, mu0, lb_h = muons.coulomb_hard_scattering(Kd, 1.0)
    for got, name in ((fCM, "cd_rock_fcm"), (screen, "cd_rock_screen"), (fspin, "cd_rock_fspin"),
                      (invlambda, "cd_rock_invlambda"), (G, "ct_rock_one"), (mu0, "hs_rock_mu0"),
                      (lb_h, "hs_rock_lbh")):
        assert same(got.cpu().numpy(), cg[name]), name
    assert same(muons.soft_scattering(Kd).cpu().numpy(), cg["ss_rock"])


@pytest.mark.gpu
def test_gpu_argument_checks():
    import torch
    from noa_b200 import dcs
    K = torch.ones(8, dtype=torch.float64, device="cuda")
    bad = torch.zeros(7, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        dcs.soft_scattering(bad, K, ELEMENTS["rock"], MUON_MASS)
    with pytest.raises(ValueError):
        dcs.coulomb_transport(torch.zeros(8, 2, dtype=torch.float64, device="cuda"),
                              torch.zeros(8, 9, dtype=torch.float64, device="cuda"), K,
                              torch.ones(3, dtype=torch.float64, device="cuda"))
    empty = torch.zeros(0, dtype=torch.float64, device="cuda")
    dcs.soft_scattering(empty, empty, ELEMENTS["rock"], MUON_MASS)      # n = 0 is a no-op
