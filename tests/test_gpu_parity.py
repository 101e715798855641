"""GPU parity tests (run on the B200): the CUDA path, called through the C ABI via the
noa::pms::dcs mirror, against the oracle on the same inputs.

Tolerance: the kernels execute the reference's IEEE operation sequence with glibc's own exp/log
algorithm, so the expected agreement is BIT FOR BIT; the asserted bar is north_star's
max-relative <= 1e-12 with exact zeros where the reference returns zero (TOL below), and the
bit-exact count is asserted separately so a regression to "merely close" is visible.
"""
import numpy as np
import pytest
import torch

from conftest import ELEMENTS, MUON_MASS
from noa_b200 import dcs, grids, physics

pytestmark = pytest.mark.gpu
TOL = 1e-12     # north_star: <= 1e-12 relative in FP64
PROC = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_parity(got, want, what, exact=True):
    got = got.detach().cpu().numpy().reshape(-1)
    want = np.asarray(want).reshape(-1)
    assert got.shape == want.shape, what
    zero = want == 0
    assert np.all(got[zero] == 0), f"{what}: non-zero where the reference is exactly 0"
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), f"{what}: NaN pattern differs"
    ok = ~zero & ~nan
    rel = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), np.finfo(np.float64).tiny)
    worst = rel.max() if rel.size else 0.0
    assert worst <= TOL, f"{what}: max relative error {worst:.3e} > {TOL}"
    if exact:
        same = np.array_equal(got, want, equal_nan=True)
        assert same, f"{what}: within {TOL} (max rel {worst:.3e}) but not bit-exact " \
                     f"({np.sum(got != want)} of {got.size} differ)"


@pytest.mark.parametrize("grid", ["A", "B", "N"])
def test_vmap_against_golden(golden, grid):
    K, q = dev(golden[grid + "_K"]), dev(golden[grid + "_q"])
    for en, el in ELEMENTS.items():
        if grid == "N" and en != "rock":
            continue
        for pr in dcs.PROCESSES:
            result = torch.zeros_like(K)
            dcs.vmap(pr)(result, K, q, el, MUON_MASS)
            assert_parity(result, golden[f"vmap_{grid}_{en}_{pr.name}"], (grid, en, pr.name))


def _same_values(got, want):
    got = got.detach().cpu().numpy().reshape(-1)
    return (got == want) | (np.isnan(got) & np.isnan(want))


def test_special_values_follow_the_reference(special):
    """Zeros, subnormals, the kinematic thresholds, q >= K, huge values, inf, NaN, negative energies
    (tests/golden/make_special_golden.py): the kernels fold their division / exp / log special
    cases into a flag and re-evaluate with the plain operations when it drops -- this is the input
    set where that second path does all the work.  Same value, inf or NaN position everywhere."""
    from conftest import SPECIAL_ELEMENTS
    import ctypes
    from noa_b200 import _lib
    K, q = dev(special["S_K"]), dev(special["S_q"])
    Kt = dev(special["ST_K"])
    lib = _lib.require_device()
    count = ctypes.c_int64(0)
    _lib.check(lib.noa_dcs_div_recomputes(ctypes.byref(count), 1))
    for en in SPECIAL_ELEMENTS:
        el = ELEMENTS[en]
        for pr in dcs.PROCESSES:
            got = dcs.map(pr)(K, q, el, MUON_MASS)
            ok = _same_values(got, special[f"vmap_S_{en}_{pr.name}"])
            assert ok.all(), (en, pr.name, special["S_K"][~ok][:4], special["S_q"][~ok][:4],
                              got.cpu().numpy()[~ok][:4], special[f"vmap_S_{en}_{pr.name}"][~ok][:4])
            for ig in (dcs.del_integrand, dcs.cel_integrand):
                res = torch.zeros_like(Kt)
                dcs.vmap_integral(dcs.recoil_integral(pr, ig))(res, Kt, 0.05, el, MUON_MASS, 180)
                want = special[f"integral_S_{en}_{pr.name}_{ig.name[:3]}_180"]
                ok = _same_values(res, want)
                assert ok.all(), (en, pr.name, ig.name, special["ST_K"][~ok],
                                  res.cpu().numpy()[~ok], want[~ok])
        # the fused forms see the same inputs
        allp = dcs.cuda.map_all(K, q, el, MUON_MASS)
        for pr in dcs.PROCESSES:
            assert _same_values(allp[pr.index], special[f"vmap_S_{en}_{pr.name}"]).all(), (en, pr)
    _lib.check(lib.noa_dcs_div_recomputes(ctypes.byref(count), 1))
    assert count.value > 0      # the plain-operation path really ran


def test_wild_inputs_follow_the_reference(port):
    """2^18 (K, q) pairs of raw bit patterns and magnitudes from the whole double range (seeded,
    conftest.wild_inputs): same value, inf or NaN position as the oracle for every process, also
    through the fused and the host-buffer entry points."""
    from conftest import wild_inputs
    K, q = wild_inputs(1 << 18)
    Kd, qd = dev(K), dev(q)
    with np.errstate(all="ignore"):
        want = [port.vmap(p, K, q, ELEMENTS["rock"], MUON_MASS, threads=8) for p in range(4)]
    for pr in dcs.PROCESSES:
        got = dcs.map(pr)(Kd, qd, ELEMENTS["rock"], MUON_MASS)
        assert _same_values(got, want[pr.index]).all(), pr.name
    allp = dcs.cuda.map_all(Kd, qd, ELEMENTS["rock"], MUON_MASS)
    for pr in dcs.PROCESSES:
        assert _same_values(allp[pr.index], want[pr.index]).all(), ("fused", pr.name)
    st = dcs.HostStager()
    try:
        out = st.map(dcs.pair_production, torch.from_numpy(K).pin_memory(),
                     torch.from_numpy(q).pin_memory(), ELEMENTS["rock"], MUON_MASS)
        assert _same_values(out, want[1]).all()
    finally:
        st.close()


def test_bremsstrahlung_electron_term_threshold(port):
    """The electron term of bremsstrahlung switches on at q < qe_max = E / (1 + m^2 / (2 me E))
    (physics.hh:137, 146).  The kernel decides that comparison from an approximation and only forms
    the quotients inside a 1e-12 band around the threshold: recoil energies ON the threshold, one
    ulp either side and at relative distances from 1e-15 to 1e-9 must follow the reference."""
    me = 0.510998910E-03
    K = 10.0 ** np.linspace(-3, 7, 4001)
    E = K + MUON_MASS
    q0 = E / (1. + 0.5 * MUON_MASS * MUON_MASS / (me * E))
    variants = [q0, np.nextafter(q0, 0.0), np.nextafter(q0, np.inf)]
    for eps in (1e-15, 3e-15, 1e-14, 1e-13, 5e-13, 9.9e-13, 1.01e-12, 2e-12, 1e-11, 1e-9):
        variants += [q0 * (1. - eps), q0 * (1. + eps)]
    Kall = np.tile(K, len(variants))
    qall = np.concatenate(variants)
    for en in ("rock", "H", "Pb"):
        got = dcs.map(dcs.bremsstrahlung)(dev(Kall), dev(qall), ELEMENTS[en], MUON_MASS)
        want = port.vmap(0, Kall, qall, ELEMENTS[en], MUON_MASS, threads=8)
        assert _same_values(got, want).all(), en
        # the threshold really is inside the sample: both sides of the switch occur
        on = port.vmap(0, K, variants[1], ELEMENTS[en], MUON_MASS)
        off = port.vmap(0, K, variants[2], ELEMENTS[en], MUON_MASS)
        assert np.any(on != off)


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 100003])
def test_vmap_ragged_sizes_against_oracle(port, n):
    K, q = grids.set_a(n)
    for pr in dcs.PROCESSES:
        got = dcs.map(pr)(dev(K), dev(q), ELEMENTS["rock"], MUON_MASS)
        assert_parity(got, port.vmap(pr.index, K, q, ELEMENTS["rock"], MUON_MASS, threads=8),
                      (n, pr.name))


def test_vmap_unaligned_views(port):
    K, q = grids.set_b(4099)
    Kd, qd = dev(K), dev(q)
    for pr in (dcs.bremsstrahlung, dcs.ionisation):
        out = torch.zeros(4100, dtype=torch.float64, device="cuda")
        view = out[1:4099 + 1 - 1]          # 8-byte but not 16-byte aligned
        Kv, qv = Kd[1:], qd[1:]
        dcs.vmap(pr)(view, Kv, qv, ELEMENTS["rock"], MUON_MASS)
        assert_parity(view, port.vmap(pr.index, K[1:], q[1:], ELEMENTS["rock"], MUON_MASS),
                      ("unaligned", pr.name))
        assert out[0] == 0 and out[-1] == 0


def test_empty_and_errors():
    e = torch.zeros(0, dtype=torch.float64, device="cuda")
    for pr in dcs.PROCESSES:
        assert dcs.map(pr)(e, e, ELEMENTS["rock"], MUON_MASS).numel() == 0
    K = torch.ones(8, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        dcs.vmap(dcs.bremsstrahlung)(K[:4], K, K, ELEMENTS["rock"], MUON_MASS)
    with pytest.raises(ValueError):
        dcs.vmap(dcs.bremsstrahlung)(K, K.float(), K, ELEMENTS["rock"], MUON_MASS)
    with pytest.raises(ValueError):
        dcs.vmap(dcs.bremsstrahlung)(K, K.cpu(), K, ELEMENTS["rock"], MUON_MASS)
    with pytest.raises(ValueError):
        dcs.vmap(dcs.bremsstrahlung)(K, K.repeat(2)[::2], K, ELEMENTS["rock"], MUON_MASS)


def test_reference_cuda_surface(golden):
    """test/unit/test-dcs-calc-cuda.cc:12-19 restated: vmap_bremsstrahlung / map_bremsstrahlung."""
    K, q = dev(golden["N_K"]), dev(golden["N_q"])
    result = torch.zeros_like(K)
    dcs.cuda.vmap_bremsstrahlung(result, K, q, physics.STANDARD_ROCK, physics.MUON_MASS)
    want = golden["vmap_N_rock_bremsstrahlung"]
    assert_parity(result, want, "vmap_bremsstrahlung")
    mapped = dcs.cuda.map_bremsstrahlung(K, q, physics.STANDARD_ROCK, physics.MUON_MASS)
    assert torch.equal(mapped, result)
    # the reference's own metric (src/noa/utils/common.hh:228-236) and threshold (1e-11)
    c, e = result.cpu().numpy(), want
    assert np.mean(np.abs((c - e) / (c + np.finfo(np.float64).tiny))) < 1e-11


def test_pair_lane_mapping_gives_identical_results(port):
    """The node-per-lane mapping of pair production (measurement library, include/
    noa_dcs_b200_probe.h) against the product's pair-per-thread kernel."""
    import ctypes
    from noa_b200 import _lib
    probe = _lib.load_probe()
    K, q = grids.set_a(50000)
    Kd, qd = dev(K), dev(q)
    a = dcs.map(dcs.pair_production)(Kd, qd, ELEMENTS["Pb"], MUON_MASS)
    b = torch.empty_like(a)
    A, I, Z = ELEMENTS["Pb"]
    vp = ctypes.c_void_p
    _lib.check(probe.noa_dcs_probe_pair_lanes_f64(
        vp(Kd.data_ptr()), vp(qd.data_ptr()), vp(b.data_ptr()), K.size, A, I, Z, MUON_MASS,
        vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert_parity(b, port.vmap(1, K, q, ELEMENTS["Pb"], MUON_MASS, threads=8), "pair lanes")


def test_fused_all_four(port):
    K, q = grids.set_a(30000)
    got = dcs.cuda.map_all(dev(K), dev(q), ELEMENTS["Fe"], MUON_MASS)
    assert got.shape == (4, 30000)
    for p in range(4):
        assert_parity(got[p], port.vmap(p, K, q, ELEMENTS["Fe"], MUON_MASS, threads=8),
                      ("all", p))


def test_water_mixture(port):
    """BASELINE config 3 semantics: sum_e w_e DCS_e, accumulated in element order from 0."""
    K, q = grids.set_b(20000)
    got = dcs.cuda.map_material(dev(K), dev(q), physics.WATER, MUON_MASS)
    for p in range(4):
        acc = np.zeros_like(K)
        for el, w in zip(physics.WATER.elements, physics.WATER.fractions):
            acc = acc + w * port.vmap(p, K, q, tuple(el), MUON_MASS, threads=8)
        assert_parity(got[p], acc, ("water", p))
    # a process subset lands in consecutive slots
    sub = dcs.cuda.map_material(dev(K), dev(q), physics.WATER, MUON_MASS,
                                processes=(dcs.pair_production, dcs.ionisation))
    assert torch.equal(sub[0], got[1]) and torch.equal(sub[1], got[3])


@pytest.mark.parametrize("min_points", [180, 1000])
def test_table_integrals_against_golden(golden, min_points):
    """test/unit/test-dcs-calc.cc:22-131 restated (DEL/CEL of every process), at the reference's
    180 nodes and at BASELINE config 4's 1000."""
    K = dev(golden["T_K"])
    for en in ("rock", "H", "Pb"):
        for pr in dcs.PROCESSES:
            for ig in (dcs.del_integrand, dcs.cel_integrand):
                result = torch.zeros_like(K)
                dcs.vmap_integral(dcs.recoil_integral(pr, ig))(
                    result, K, dcs.X_FRACTION, ELEMENTS[en], MUON_MASS, min_points)
                key = f"integral_{en}_{pr.name}_{ig.name[:3]}_{min_points}"
                assert_parity(result, golden[key], key)


def test_fused_tables_equal_single_columns(golden):
    K = dev(golden["T_K"])
    del_t, cel_t = dcs.cuda.tables(K, 0.05, ELEMENTS["rock"], MUON_MASS, 1000)
    for pr in dcs.PROCESSES:
        assert_parity(del_t[pr.index], golden[f"integral_rock_{pr.name}_del_1000"], pr.name)
        assert_parity(cel_t[pr.index], golden[f"integral_rock_{pr.name}_cel_1000"], pr.name)
    # subset: untouched rows stay zero
    d2, c2 = dcs.cuda.tables(K, 0.05, ELEMENTS["rock"], MUON_MASS, 1000,
                             processes=(dcs.bremsstrahlung,))
    assert torch.equal(d2[0], del_t[0]) and float(d2[1:].abs().sum()) == 0.0


def test_material_tables_mix_element_tables(port):
    """Water = H + O: table = sum_e t_e * w_e in composition order (pumas.c:8054-8078), with the
    element tables themselves bit-identical to the oracle's recoil integrals."""
    from noa_b200 import WATER
    K = grids.table_energies(97, -2.0, 6.0)
    table, parts = dcs.cuda.material_tables(dev(K), 0.05, WATER, MUON_MASS, 180)
    assert table.shape == (2, 4, 97) and parts.shape == (2, 2, 4, 97)
    want = np.zeros((2, 4, 97))
    for e, (el, w) in enumerate(zip(WATER.elements, WATER.fractions)):
        for ig in range(2):
            for p in range(4):
                t = port.vmap_integral(p, ig, K, 0.05, 180, tuple(el), MUON_MASS, threads=8)
                assert_parity(parts[e, ig, p], t, ("element table", e, ig, p))
                want[ig, p] += t * w
    assert np.array_equal(table.cpu().numpy(), want)
    # a process subset leaves the other rows at exactly zero
    t2, _ = dcs.cuda.material_tables(dev(K), 0.05, WATER, MUON_MASS, 180,
                                     processes=(dcs.photonuclear,))
    assert torch.equal(t2[:, 2], table[:, 2]) and float(t2[:, [0, 1, 3]].abs().sum()) == 0.0


@pytest.mark.parametrize("n_rows", [1, 2, 3, 255, 1001])
def test_table_row_pairing_and_thresholds(port, n_rows):
    """The two cheap processes put two rows into one CTA pass; ionisation switches between the
    closed form and the quadrature at K = 0.5 (m - me)^2 / me = 10.8 GeV: odd row counts, a single
    row, and a grid that straddles the switch (so that a pass holds one row of each kind), all
    four processes, both integrands, against the oracle."""
    K = grids.table_energies(n_rows, 0.9, 1.2) if n_rows > 1 else np.array([10.9])
    d, c = dcs.cuda.tables(dev(K), 0.05, ELEMENTS["rock"], MUON_MASS, 1000)
    for pr in dcs.PROCESSES:
        for ig, got in ((0, d), (1, c)):
            want = port.vmap_integral(pr.index, ig, K, 0.05, 1000, ELEMENTS["rock"], MUON_MASS,
                                      threads=8)
            assert_parity(got[pr.index], want, ("pairing", n_rows, pr.name, ig))


def test_table_config4_sample_against_oracle(port):
    """BASELINE config 4's own grid (10^4 energies, 1e-2 .. 1e6 GeV, 1000 points): the full GPU
    table, every 20th energy of it against the oracle (bench.py checks all 8 x 10^4 values against
    the compiled reference on the B200 box)."""
    K = grids.table_energies(10000)
    d, c = dcs.cuda.tables(dev(K), 0.05, ELEMENTS["rock"], MUON_MASS, 1000)
    idx = np.arange(0, K.size, 20)
    sel = torch.from_numpy(idx).cuda()
    for pr in dcs.PROCESSES:
        for ig, got in ((0, d), (1, c)):
            want = port.vmap_integral(pr.index, ig, K[idx], 0.05, 1000, ELEMENTS["rock"],
                                      MUON_MASS, threads=16)
            assert_parity(got[pr.index][sel], want, ("config 4", pr.name, ig))
    # a second build gives the same bits (no dependence on CTA scheduling)
    d2, c2 = dcs.cuda.tables(dev(K), 0.05, ELEMENTS["rock"], MUON_MASS, 1000)
    assert torch.equal(d, d2) and torch.equal(c, c2)


def test_table_odd_node_counts(port):
    """min_points not a multiple of 6 and larger than one shared-memory pass (1536 nodes)."""
    K = grids.table_energies(24, -1.0, 5.0)
    for mp in (1, 7, 1537, 4000):
        for pr in (dcs.bremsstrahlung, dcs.ionisation, dcs.pair_production):
            result = torch.zeros(24, dtype=torch.float64, device="cuda")
            dcs.vmap_integral(dcs.recoil_integral(pr, dcs.cel_integrand))(
                result, dev(K), 0.05, ELEMENTS["rock"], MUON_MASS, mp)
            want = port.vmap_integral(pr.index, 1, K, 0.05, mp, ELEMENTS["rock"], MUON_MASS,
                                      threads=8)
            assert_parity(result, want, (mp, pr.name))


@pytest.mark.parametrize("min_points", [1, 7, 31, 180, 257, 1000, 1537, 4000])
def test_flat_tables_against_oracle_and_rows_form(port, min_points):
    """The workspace ("flat") form of the table build -- 32-node units popped by warps, terms
    through the workspace, summation kernel -- against the oracle's recoil integrals and against
    the row-per-CTA form, at node counts below / at / above a unit, a summation stage (256) and a
    shared-memory pass (1536), on a ragged number of rows that spans the ionisation closed-form
    threshold and the kinematic thresholds of pair production and photonuclear."""
    K = grids.table_energies(37, -2.0, 6.0)
    Kd = dev(K)
    d_flat, c_flat = dcs.cuda.tables(Kd, 0.05, ELEMENTS["rock"], MUON_MASS, min_points, flat=True)
    d_rows, c_rows = dcs.cuda.tables(Kd, 0.05, ELEMENTS["rock"], MUON_MASS, min_points, flat=False)
    assert torch.equal(d_flat, d_rows) and torch.equal(c_flat, c_rows)
    for pr in dcs.PROCESSES:
        for ig, got in ((0, d_flat), (1, c_flat)):
            want = port.vmap_integral(pr.index, ig, K, 0.05, min_points, ELEMENTS["rock"],
                                      MUON_MASS, threads=8)
            assert_parity(got[pr.index], want, (min_points, pr.name, ig))


@pytest.mark.parametrize("mask", [1, 2, 4, 8, 9, 6, 10, 7, 14])
def test_flat_tables_process_subsets(mask):
    """Every way the flat build groups its launches (bremsstrahlung + ionisation fused or alone,
    heavy processes present or not): requested rows equal the full build's, the others are zero --
    also when the output held something else before."""
    K = dev(grids.table_energies(203, -2.0, 6.0))
    full_d, full_c = dcs.cuda.tables(K, 0.05, ELEMENTS["Pb"], MUON_MASS, 180, flat=True)
    procs = tuple(pr for pr in dcs.PROCESSES if (mask >> pr.index) & 1)
    out = (torch.full((4, 203), 7.0, dtype=torch.float64, device="cuda"),
           torch.full((4, 203), -3.0, dtype=torch.float64, device="cuda"))
    d, c = dcs.cuda.tables(K, 0.05, ELEMENTS["Pb"], MUON_MASS, 180, processes=procs, out=out,
                           flat=True)
    for pr in dcs.PROCESSES:
        if (mask >> pr.index) & 1:
            assert torch.equal(d[pr.index], full_d[pr.index]), (mask, pr.name)
            assert torch.equal(c[pr.index], full_c[pr.index]), (mask, pr.name)
        else:
            assert float(d[pr.index].abs().sum()) == 0.0 and float(c[pr.index].abs().sum()) == 0.0


@pytest.mark.parametrize("mask", [15, 13, 9, 5])
def test_flat_tables_with_many_rows(port, mask):
    """From 4 096 rows on bremsstrahlung and ionisation ride on the photonuclear kernel's nodes
    (one exp for three integrands): same bits as the row-per-CTA form and as the oracle, for every
    process subset that takes or just misses that path."""
    n = 4500
    K = grids.table_energies(n, -2.0, 6.0)
    Kd = dev(K)
    procs = tuple(pr for pr in dcs.PROCESSES if (mask >> pr.index) & 1)
    d_flat, c_flat = dcs.cuda.tables(Kd, 0.05, ELEMENTS["Fe"], MUON_MASS, 31, processes=procs,
                                     flat=True)
    d_rows, c_rows = dcs.cuda.tables(Kd, 0.05, ELEMENTS["Fe"], MUON_MASS, 31, processes=procs,
                                     flat=False)
    assert torch.equal(d_flat, d_rows) and torch.equal(c_flat, c_rows)
    idx = np.arange(0, n, 53)
    for pr in procs:
        want = port.vmap_integral(pr.index, 1, K[idx], 0.05, 31, ELEMENTS["Fe"], MUON_MASS,
                                  threads=8)
        assert_parity(c_flat[pr.index][torch.from_numpy(idx).cuda()], want, (mask, pr.name))


def test_table_workspace_contract():
    """C ABI: noa_dcs_table_workspace_doubles sizes the workspace; a NULL or short workspace makes
    noa_dcs_table_ws_f64 fall back to the row-per-CTA launches (same bits); the exchange form
    refuses to run without one."""
    import ctypes
    from noa_b200 import _lib
    lib = _lib.require_device()
    vp = ctypes.c_void_p
    n, mp = 100, 180
    need = int(lib.noa_dcs_table_workspace_doubles(n, mp))
    assert need >= 8 * n * 180 and lib.noa_dcs_table_workspace_doubles(0, mp) == 0
    K = dev(grids.table_energies(n, -2.0, 6.0))
    ws = torch.empty(need, dtype=torch.float64, device="cuda")
    stream = vp(torch.cuda.current_stream().cuda_stream)
    tabs = []
    for doubles in (need, need - 1, 0):
        t = torch.full((2, 4, n), 5.0, dtype=torch.float64, device="cuda")
        _lib.check(lib.noa_dcs_table_ws_f64(15, vp(K.data_ptr()), n, 0.05, mp, 22., 0.1364e-6, 11,
                                            MUON_MASS, vp(t.data_ptr()),
                                            vp(t.data_ptr() + 4 * n * 8),
                                            vp(ws.data_ptr()) if doubles else None, doubles, stream))
        tabs.append(t)
    torch.cuda.synchronize()
    assert torch.equal(tabs[0], tabs[1]) and torch.equal(tabs[0], tabs[2])
    # exchange form, this GPU as its only peer, no scratch: invalid argument, nothing launched
    flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    sync = torch.zeros(8, dtype=torch.int32, device="cuda")
    t = tabs[0]
    dl, cl, fl = (vp * 1)(t.data_ptr()), (vp * 1)(t.data_ptr() + 4 * n * 8), (vp * 1)(flags.data_ptr())
    rc = lib.noa_dcs_table_exchange_f64(15, vp(K.data_ptr()), n, 0.05, mp, 22., 0.1364e-6, 11,
                                        MUON_MASS, 1, 0, dl, cl, fl, None, None, None, vp(sync.data_ptr()), None, 0,
                                        1, n, 0, 1, 5.0, stream)
    assert rc == -1      # NOA_DCS_EINVAL
    # ... and with one: the same table
    t2 = torch.zeros((2, 4, n), dtype=torch.float64, device="cuda")
    dl2, cl2 = (vp * 1)(t2.data_ptr()), (vp * 1)(t2.data_ptr() + 4 * n * 8)
    _lib.check(lib.noa_dcs_table_exchange_f64(15, vp(K.data_ptr()), n, 0.05, mp, 22., 0.1364e-6,
                                              11, MUON_MASS, 1, 0, dl2, cl2, fl, None, None, None,
                                              vp(sync.data_ptr()), vp(ws.data_ptr()), need, 1, n,
                                              0, 1, 5.0, stream))
    torch.cuda.synchronize()
    assert torch.equal(t2, tabs[0])


def test_full_size_properties():
    """BASELINE config 2 size (2^22 pairs): size-independent checks -- determinism, agreement of a
    strided sample with the oracle is covered above; here: repeatability, slice consistency
    (vmap of a slice == slice of vmap) and finiteness on the in-range grid."""
    n = 1 << 22
    K, q = grids.set_b(n)
    Kd, qd = dev(K), dev(q)
    a = dcs.map(dcs.pair_production)(Kd, qd, ELEMENTS["rock"], MUON_MASS)
    b = dcs.map(dcs.pair_production)(Kd, qd, ELEMENTS["rock"], MUON_MASS)
    assert torch.equal(a, b)
    assert bool(torch.isfinite(a).all()) and bool((a > 0).all())
    lo, hi = 1234567, 1234567 + 65536
    c = dcs.map(dcs.pair_production)(Kd[lo:hi].contiguous(), qd[lo:hi].contiguous(),
                                     ELEMENTS["rock"], MUON_MASS)
    assert torch.equal(c, a[lo:hi])


def test_full_size_sample_against_oracle(port):
    n = 1 << 22
    K, q = grids.set_b(n)
    Kd, qd = dev(K), dev(q)
    idx = np.arange(0, n, 509)
    for pr in dcs.PROCESSES:
        full = dcs.map(pr)(Kd, qd, ELEMENTS["rock"], MUON_MASS)
        want = port.vmap(pr.index, K[idx], q[idx], ELEMENTS["rock"], MUON_MASS, threads=8)
        assert_parity(full[torch.from_numpy(idx).cuda()], want, ("2^22 sample", pr.name))


def test_tau_projectile(port):
    """A projectile other than the muon (tau, physics.hh:59): every mass-dependent invariant and
    staged reciprocal, element-wise and through the table integrals."""
    tau = 1.77682
    K, q = grids.set_a(1 << 15)
    Kt = grids.table_energies(48, -2.0, 6.0)
    for en in ("rock", "H", "Pb"):
        for pr in dcs.PROCESSES:
            got = dcs.map(pr)(dev(K), dev(q), ELEMENTS[en], tau)
            assert_parity(got, port.vmap(pr.index, K, q, ELEMENTS[en], tau, threads=8),
                          ("tau", en, pr.name))
        d, c = dcs.cuda.tables(dev(Kt), dcs.X_FRACTION, ELEMENTS[en], tau, 180)
        for pr in dcs.PROCESSES:
            assert_parity(d[pr.index], port.vmap_integral(pr.index, 0, Kt, 0.05, 180, ELEMENTS[en],
                                                          tau), ("tau DEL", en, pr.name))
            assert_parity(c[pr.index], port.vmap_integral(pr.index, 1, Kt, 0.05, 180, ELEMENTS[en],
                                                          tau), ("tau CEL", en, pr.name))


def test_host_stager_roundtrip(port):
    K, q = grids.set_a(300001)
    Kh = torch.from_numpy(K).pin_memory()
    qh = torch.from_numpy(q).pin_memory()
    st = dcs.HostStager(chunk_pairs=1 << 16, n_slots=3)
    try:
        for pr in (dcs.bremsstrahlung, dcs.pair_production):
            want = port.vmap(pr.index, K, q, ELEMENTS["rock"], MUON_MASS, threads=8)
            out = st.map(pr, Kh, qh, ELEMENTS["rock"], MUON_MASS)      # pinned: in-place kernel
            assert not out.is_cuda and out.is_pinned()
            assert_parity(out, want, ("host pinned", pr.name))
            # pageable tensors and an unaligned pinned view go through the copy pipeline / the
            # scalar-load kernel
            out = st.map(pr, torch.from_numpy(K), torch.from_numpy(q), ELEMENTS["rock"], MUON_MASS)
            assert_parity(out, want, ("host pageable", pr.name))
            out = st.map(pr, Kh[1:], qh[1:], ELEMENTS["rock"], MUON_MASS)
            assert_parity(out, want[1:], ("host pinned, 8-byte aligned view", pr.name))
            st.zero_copy = False
            out = st.map(pr, Kh, qh, ELEMENTS["rock"], MUON_MASS)
            st.zero_copy = True
            assert_parity(out, want, ("host staged", pr.name))
        allp = st.map(None, Kh, qh, ELEMENTS["rock"], MUON_MASS)
        assert allp.shape == (4, 300001)
        assert_parity(allp[2], port.vmap(2, K, q, ELEMENTS["rock"], MUON_MASS, threads=8),
                      "host all/photonuclear")
    finally:
        st.close()


def test_runs_on_a_side_stream(port):
    K, q = grids.set_b(10000)
    Kd, qd = dev(K), dev(q)
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        out = dcs.map(dcs.photonuclear)(Kd, qd, ELEMENTS["rock"], MUON_MASS)
    s.synchronize()
    assert_parity(out, port.vmap(2, K, q, ELEMENTS["rock"], MUON_MASS, threads=8), "side stream")


def test_cpp_libtorch_boundary_via_pybind(golden):
    """The C++ API (noa::pms::dcs::cuda::*, csrc/torch_api.cc) through the muon_dcs pybind module
    that mirrors docs/pms/muon_dcs.{cc,cu}."""
    from noa_b200 import muons
    K, q = dev(golden["N_K"]), dev(golden["N_q"])
    for name in PROC:
        got = getattr(muons, name)(K, q)
        assert_parity(got, golden[f"vmap_N_rock_{name}"], ("muons", name))
    allp = muons.all_processes(K, q)
    assert allp.shape == (4, K.numel())
    assert_parity(allp[1], golden["vmap_N_rock_pair_production"], "muons.all_processes")
    Kt = dev(golden["T_K"])
    t = muons.tables(Kt, 0.05, 180)
    assert t.shape == (2, 4, Kt.numel())
    assert_parity(t[0, 2], golden["integral_rock_photonuclear_del_180"], "muons.tables del")
    assert_parity(t[1, 3], golden["integral_rock_ionisation_cel_180"], "muons.tables cel")
    col = muons.recoil_integral(0, 1, Kt, 0.05, 1000)
    assert_parity(col, golden["integral_rock_bremsstrahlung_cel_1000"], "muons.recoil_integral")
    with pytest.raises(RuntimeError):
        muons.bremsstrahlung(K.float(), q.float())
    with pytest.raises(RuntimeError):
        muons.bremsstrahlung(K, q.cpu())                 # tensors on different devices


def test_cpp_reference_cpu_call_sites(golden):
    """The reference's CPU call expressions -- dcs::vmap(dcs::pair_production)(result, K, q, ...),
    dcs::vmap_integral(dcs::recoil_integral(f, g))(result, K, xlow, ...), dcs::map / pmap / pvmap
    (test/unit/test-dcs-calc.cc:22-131, docs/pms/muon_dcs.cc:8-27) -- compiled unchanged against
    include/noa_b200/pms_dcs.hh, on CPU tensors (pageable and pinned) and on CUDA tensors."""
    from noa_b200 import muons
    Kn, qn = golden["N_K"], golden["N_q"]
    Kt = golden["T_K"]
    variants = {
        "pageable": (torch.from_numpy(Kn.copy()), torch.from_numpy(qn.copy())),
        "pinned": (torch.from_numpy(Kn.copy()).pin_memory(), torch.from_numpy(qn.copy()).pin_memory()),
        "cuda": (dev(Kn), dev(qn)),
    }
    for label, (K, q) in variants.items():
        out = muons.reference_call_sites(K, q)
        assert len(out) == 15
        for t in out:
            assert t.device == K.device, label
        for i, name in enumerate(PROC):
            assert_parity(out[i], golden[f"vmap_N_rock_{name}"], (label, "vmap", name))
        for t in out[12:]:
            assert_parity(t, golden["vmap_N_rock_pair_production"], (label, "map/pmap/pvmap"))
    # the integral columns are pinned by the golden table energies
    for label, K in (("pageable", torch.from_numpy(Kt.copy())), ("cuda", dev(Kt))):
        out = muons.reference_call_sites(K, K * 0.0505)
        i = 4
        for name in PROC:
            for ig in ("del", "cel"):
                assert_parity(out[i], golden[f"integral_rock_{name}_{ig}_180"], (label, name, ig))
                i += 1
    # CPU tensors in, CPU tensors out for the table builders too
    t = muons.tables(torch.from_numpy(Kt.copy()), 0.05, 180)
    assert not t.is_cuda
    assert_parity(t[0, 2], golden["integral_rock_photonuclear_del_180"], "tables on CPU tensors")
    allp = muons.all_processes(variants["pageable"][0], variants["pageable"][1])
    assert not allp.is_cuda
    assert_parity(allp[3], golden["vmap_N_rock_ionisation"], "all_processes on CPU tensors")


def test_reference_benchmark_cases_cli():
    """benchmark/measure_dcs_calc_cuda: every case name of the reference's measure-dcs-calc harness
    that has a GPU form prints one Google-Benchmark-style line."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "noa_b200", "measure_dcs_calc_cuda")
    if not os.path.exists(exe):
        pytest.skip("benchmark executable not built")
    r = subprocess.run([exe, "--benchmark_min_time=0.02"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    for case in ("BremsstrahlungVectorisedCUDA", "BremsstrahlungVectorisedLargeCUDA",
                 "PairProductionVectorisedCUDA", "PhotonuclearVectorisedLargeCUDA",
                 "DELIonisationVectorisedCUDA", "CELPairProductionVectorisedCUDA",
                 "CoulombHardScatteringCUDA", "CoulombSoftScatteringCUDA"):
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("DCSBenchmark/" + case + " ")]
        assert len(lines) == 1, case
        assert float(lines[0].split()[1]) > 0
