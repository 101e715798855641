"""Per-material table assembly: the PUMAS steps
(src/noa/3rdparty/_pumas/pumas.c:8054-8111, 10768-10808, 10816-10881, 10901-10955) over NOA's DCS.

The golden fixture (tests/golden/material_golden.npz, tests/golden/make_material_golden.py) was
produced by oracle/material_oracle.c driving the COMPILED REFERENCE's scalar DCS and quadrature.
CPU tests: the C port reproduces it bit for bit (and the compiled reference, where present, still
does).  GPU tests: the CUDA assembly reproduces it bit for bit -- including every branch of the
threshold search, which only works because each probe DCS value is bit-identical."""
import os

import numpy as np
import pytest

from conftest import ELEMENTS, GOLDEN_DIR, MUON_MASS

H, O, ROCK, FE, PB = (ELEMENTS[k] for k in ("H", "O", "rock", "Fe", "Pb"))
MATERIALS = {"water": ((H, O), (0.111894, 0.888106)), "rock": ((ROCK,), (1.0,)),
             "lead_iron": ((PB, FE), (0.5, 0.5))}
KEYS = ("elem", "cs", "cel", "straggling", "csf", "cs_total", "xt")
PROC = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")


@pytest.fixture(scope="module")
def mgolden():
    return np.load(os.path.join(GOLDEN_DIR, "material_golden.npz"))


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("name", ["water", "rock", "lead_iron"])
def test_port_assembly_matches_golden(port, reference, mgolden, name):
    import oracle
    elements, fractions = MATERIALS[name]
    for checker in (port, reference):
        if checker is None:
            continue
        r = oracle.material_assembly(checker, elements, fractions, MUON_MASS, mgolden["M_K"], 0.05,
                                     180, threads=8)
        assert r["rc"] == 0
        for key in KEYS:
            assert _same(r[key], mgolden[f"{name}_{key}"]), (checker.kind, name, key)
        assert r["kt"] == mgolden[f"{name}_kt"][0] and r["it"] == mgolden[f"{name}_it"][0]


def test_assembly_golden_is_physically_sane(mgolden):
    """Not a parity check: guards the fixture itself (monotone cumulative fractions ending at 1,
    thresholds inside [cutoff, 1], cs / cel rows equal to the mass-fraction mix of the element
    tables, straggling only from ionisation)."""
    for name, (elements, fractions) in MATERIALS.items():
        csf, xt, elem = (mgolden[f"{name}_{k}"] for k in ("csf", "xt", "elem"))
        ne, n = csf.shape[0], csf.shape[-1]
        flat = csf.reshape(ne * 4, n)
        live = mgolden[f"{name}_cs"].sum(axis=0) > 0
        assert np.all(np.diff(flat[:, live], axis=0) >= 0) and np.all(flat[-1, live] == 1.0)
        assert np.all((xt >= 0.05) & (xt <= 1.0))
        assert not elem[:, 2, :3].any() and elem[:, 2, 3].any()
        mix = sum(elem[e, 0] * w for e, w in enumerate(fractions))
        assert np.allclose(mix, mgolden[f"{name}_cs"], rtol=1e-14, atol=0)


def test_port_integral_modes_match_golden(port, mgolden):
    import oracle
    for p, pn in enumerate(PROC):
        for mode in (0, 1, 2):
            for tag, (lo, hi) in (("cut", (1E-06, 0.05)), ("mid", (0.01, 0.5))):
                got = oracle.integral_scalar(port, p, mode, mgolden["M_K"], lo, hi, ROCK, MUON_MASS,
                                             180)
                assert _same(got, mgolden[f"mode_{pn}_{mode}_{tag}"]), (pn, mode, tag)


# ---- GPU -------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["water", "rock", "lead_iron", "water2"])
def test_gpu_assembly_matches_golden(mgolden, name):
    import torch
    from noa_b200 import dcs, physics
    elements, fractions = MATERIALS["water" if name == "water2" else name]
    material = physics.Material(name, tuple(physics.AtomicElement(*e) for e in elements), fractions)
    K = torch.from_numpy(mgolden["M2_K" if name == "water2" else "M_K"]).cuda()
    out = dcs.cuda.material_assembly(K, 0.05, material, MUON_MASS, 180)
    torch.cuda.synchronize()
    for key in KEYS:
        got, want = out[key].cpu().numpy(), mgolden[f"{name}_{key}"]
        assert _same(got, want), (name, key, int((got != want).sum()))
    assert out["kt"].item() == mgolden[f"{name}_kt"][0]
    assert out["it"].item() == mgolden[f"{name}_it"][0]


@pytest.mark.gpu
def test_gpu_integral_modes_match_golden(mgolden):
    import torch
    from noa_b200 import dcs
    K = torch.from_numpy(mgolden["M_K"]).cuda()
    for pr in dcs.PROCESSES:
        for mode in (0, 1, 2):
            for tag, (lo, hi) in (("cut", (1E-06, 0.05)), ("mid", (0.01, 0.5))):
                res = torch.zeros_like(K)
                dcs.cuda.vmap_integral_mode(res, K, pr, mode, lo, hi, ROCK, MUON_MASS, 180)
                want = mgolden[f"mode_{pr.name}_{mode}_{tag}"]
                got = res.cpu().numpy()
                assert _same(got, want), (pr.name, mode, tag, int((got != want).sum()))
    # modes 0 / 1 with x_high = 1 are dcs::recoil_integral itself (closed forms included)
    for pr in dcs.PROCESSES:
        for mode, ig in ((0, dcs.del_integrand), (1, dcs.cel_integrand)):
            a, b = torch.zeros_like(K), torch.zeros_like(K)
            dcs.cuda.vmap_integral_mode(a, K, pr, mode, 0.05, 1.0, ROCK, MUON_MASS, 180)
            dcs.vmap_integral(dcs.recoil_integral(pr, ig))(b, K, 0.05, ROCK, MUON_MASS, 180)
            assert torch.equal(a, b)


@pytest.mark.gpu
def test_gpu_assembly_large_grid_against_port(port):
    """10^3 energies, water: every output against the C port (threads), bit for bit."""
    import oracle
    import torch
    from noa_b200 import dcs, grids, physics
    K = grids.table_energies(1000)
    out = dcs.cuda.material_assembly(torch.from_numpy(K).cuda(), 0.05, physics.WATER, MUON_MASS, 180)
    want = oracle.material_assembly(port, [tuple(e) for e in physics.WATER.elements],
                                    physics.WATER.fractions, MUON_MASS, K, 0.05, 180, threads=16)
    for key in KEYS:
        assert _same(out[key].cpu().numpy(), want[key]), key
    assert out["kt"].item() == want["kt"] and out["it"].item() == want["it"]
