"""CPU tests (run everywhere): the C oracle against (1) the committed golden vectors generated
from the compiled reference, (2) the compiled reference itself where oracle/_ref exists,
(3) the values the reference prints in docs/pms/muon_dcs_calc.ipynb and SURVEY.md's 17-digit KATs."""
import numpy as np
import pytest

from conftest import ELEMENTS, MUON_MASS
from noa_b200 import grids

PROC = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")


def test_port_matches_golden_vmap_bit_for_bit(port, golden):
    for g in "ABN":
        K, q = golden[g + "_K"], golden[g + "_q"]
        for en, el in ELEMENTS.items():
            if g == "N" and en != "rock":
                continue
            for p, pn in enumerate(PROC):
                want = golden[f"vmap_{g}_{en}_{pn}"]
                got = port.vmap(p, K, q, el, MUON_MASS)
                assert np.array_equal(got, want, equal_nan=True), (g, en, pn)


def test_port_matches_golden_integrals_bit_for_bit(port, golden):
    K = golden["T_K"]
    for en in ("rock", "H", "Pb"):
        for p, pn in enumerate(PROC):
            for ig, ign in enumerate(("del", "cel")):
                for mp in (180, 1000):
                    want = golden[f"integral_{en}_{pn}_{ign}_{mp}"]
                    got = port.vmap_integral(p, ig, K, 0.05, mp, ELEMENTS[en], MUON_MASS, threads=4)
                    assert np.array_equal(got, want), (en, pn, ign, mp)


def test_port_matches_compiled_reference(port, reference):
    if reference is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    K, q = grids.set_a(1 << 15)
    for p in range(4):
        for el in (ELEMENTS["rock"], ELEMENTS["H"]):
            a = port.vmap(p, K, q, el, MUON_MASS, threads=4)
            b = reference.vmap(p, K, q, el, MUON_MASS, threads=4)
            assert np.array_equal(a, b, equal_nan=True)
    # tau projectile as well
    for p in range(4):
        a = port.vmap(p, K[:4096], q[:4096], ELEMENTS["rock"], 1.77682)
        b = reference.vmap(p, K[:4096], q[:4096], ELEMENTS["rock"], 1.77682)
        assert np.array_equal(a, b, equal_nan=True)
        Kt = grids.table_energies(24, -2.0, 6.0)
        for ig in (0, 1):
            a = port.vmap_integral(p, ig, Kt, 0.05, 180, ELEMENTS["rock"], 1.77682)
            b = reference.vmap_integral(p, ig, Kt, 0.05, 180, ELEMENTS["rock"], 1.77682)
            assert np.array_equal(a, b, equal_nan=True), (p, ig)


def test_notebook_printed_values(port):
    """docs/pms/muon_dcs_calc.ipynb:325,535,670,741 -- first five outputs, 5 significant digits."""
    K, q = grids.notebook_grid()
    printed = {
        0: [3.5293e-04, 3.9395e-06, 4.0777e-06, 4.1341e-06, 4.1650e-06],
        1: [0.0, 6.5366e-06, 7.3699e-06, 7.7919e-06, 8.0572e-06],
        2: [0.0, 2.2912e-06, 2.1304e-06, 2.0719e-06, 2.0427e-06],
        3: [0.0, 3.0168e-05, 1.5300e-05, 1.0284e-05, 7.7575e-06],
    }
    for p, want in printed.items():
        got = port.vmap(p, K[:5], q[:5], ELEMENTS["rock"], MUON_MASS)
        for g, w in zip(got, want):
            if w == 0.0:
                assert g == 0.0
            else:
                assert abs(g - w) / w < 6e-5, (p, g, w)


def test_survey_known_answers(port):
    """17-digit values of the compiled reference on set B, i = 12345."""
    K, q = grids.set_b(1 << 20, 12345, 1)
    assert K[0] == 111.45366265614962 and q[0] == 43.828980924642451
    kat = {
        "rock": [2.9528731797081247e-07, 1.0338456944588468e-08, 8.8768162121671727e-08,
                 2.9904149569829071e-07],
        "Pb": [1.4505524840424877e-06, 5.2506195749816527e-08, 7.348377024835874e-08,
               2.3669307574575136e-07],
        "H": [9.9164874904784469e-08, 3.5172877469186998e-09, 1.158170741919867e-07,
              5.9292454783045638e-07],
    }
    for en, vals in kat.items():
        for p, w in enumerate(vals):
            g = port.vmap(p, K, q, ELEMENTS[en], MUON_MASS)[0]
            assert abs(g - w) <= 1e-14 * abs(w), (en, p, g, w)
    # table integrals, standard rock, xlow 0.05, min_points 1000, K = 1e-2 .. 1e5 (decades)
    Kt = 10.0 ** np.arange(-2, 6)
    del_brems = [2.1080445272917965e-07, 1.186516641920316e-07, 1.5886668575724013e-07,
                 2.8151889557958887e-07, 4.1580955676585812e-07, 5.0472251913573816e-07,
                 5.387177842486199e-07, 5.4708075371438513e-07]
    cel_photo = [0, 0, 0, 3.2621631937832171e-07, 3.0607374533865346e-06, 3.1839946749936191e-05,
                 0.0003911171435036197, 0.0050238694108667879]
    del_ion = [0, 0, 2.479932517905856e-05, 1.0343667876283779e-05, 1.3019430570974147e-06,
               1.3863024663265029e-07, 1.4628307850766108e-08, 1.55255892065722e-09]
    cel_pair = [0, 0, 1.6478524396030436e-09, 7.3649668079363759e-08, 1.549510044625862e-06,
                2.2206831834918222e-05, 0.00025048725697641614, 0.002569236863857328]
    for (p, ig, want) in ((0, 0, del_brems), (2, 1, cel_photo), (3, 0, del_ion), (1, 1, cel_pair)):
        got = port.vmap_integral(p, ig, Kt, 0.05, 1000, ELEMENTS["rock"], MUON_MASS)
        for g, w in zip(got, want):
            assert (g == 0.0) if w == 0 else abs(g - w) <= 1e-14 * abs(w), (p, ig, g, w)


def test_edge_cases(port):
    el = ELEMENTS["rock"]
    empty = np.zeros(0)
    for p in range(4):
        assert port.vmap(p, empty, empty, el, MUON_MASS).size == 0
    # kinematic exits give exact zeros (dcs.hh:151-156, 357-359, 374, 420-424)
    K = np.array([10.0, 10.0, 0.5, 1e3, 1e3])
    q = np.array([1e-3, 11.0, 0.2, 0.5, 1.5])
    assert port.vmap(1, K, q, el, MUON_MASS)[0] == 0.0      # q <= 4 me
    assert port.vmap(1, K, q, el, MUON_MASS)[1] == 0.0      # q beyond the upper bound
    assert port.vmap(2, K, q, el, MUON_MASS)[3] == 0.0      # q < 1 GeV
    assert port.vmap(2, K, q, el, MUON_MASS)[4] == 0.0      # q < 2e-3 K
    assert port.vmap(3, K, q, el, MUON_MASS)[1] == 0.0      # q > Wmax
    # NaN propagates through bremsstrahlung (no guard, physics.hh:135-152)
    assert np.isnan(port.vmap(0, np.array([np.nan]), np.array([1.0]), el, MUON_MASS)[0])


def test_port_matches_reference_on_special_values(port, special):
    """Inputs outside anything physical: the reference does not validate, so its 0 / NaN / inf /
    garbage there is the contract (same NaN positions; NaN sign and payload are not compared)."""
    from conftest import SPECIAL_ELEMENTS
    K, q = special["S_K"], special["S_q"]
    with np.errstate(all="ignore"):
        for en in SPECIAL_ELEMENTS:
            for p, pn in enumerate(PROC):
                got = port.vmap(p, K, q, ELEMENTS[en], MUON_MASS)
                want = special[f"vmap_S_{en}_{pn}"]
                bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
                assert not bad.any(), (en, pn, K[bad][:4], q[bad][:4], got[bad][:4], want[bad][:4])
                for ig, ign in enumerate(("del", "cel")):
                    got = port.vmap_integral(p, ig, special["ST_K"], 0.05, 180, ELEMENTS[en],
                                             MUON_MASS)
                    want = special[f"integral_S_{en}_{pn}_{ign}_180"]
                    bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
                    assert not bad.any(), (en, pn, ign, special["ST_K"][bad], got[bad], want[bad])


def test_port_matches_compiled_reference_on_wild_inputs(port, reference):
    if reference is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    from conftest import wild_inputs
    K, q = wild_inputs(1 << 16)
    with np.errstate(all="ignore"):
        for p in range(4):
            for el in (ELEMENTS["rock"], ELEMENTS["H"]):
                a = port.vmap(p, K, q, el, MUON_MASS, threads=4)
                b = reference.vmap(p, K, q, el, MUON_MASS, threads=4)
                assert ((a == b) | (np.isnan(a) & np.isnan(b))).all(), p


def test_port_matches_compiled_reference_at_config_1_size(port, reference):
    """BASELINE config 1 size (2^20 pairs) on both synthetic sets, standard rock: the restatement
    and the compiled reference agree on every one of the 8 x 2^20 values."""
    if reference is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    n = 1 << 20
    for K, q in (grids.set_a(n), grids.set_b(n)):
        for p in range(4):
            a = port.vmap(p, K, q, ELEMENTS["rock"], MUON_MASS, threads=8)
            b = reference.vmap(p, K, q, ELEMENTS["rock"], MUON_MASS, threads=8)
            assert np.array_equal(a, b, equal_nan=True), p
