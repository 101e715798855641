#!/usr/bin/env python
"""Generate tests/golden/dcs_golden.npz from the COMPILED REFERENCE (oracle/_ref/libnoa_ref.so =
the unmodified headers under /root/reference/src, see oracle/Makefile).  Run in the build
container (the only place /root/reference exists); the .npz is committed and travels to the GPU
box, where neither /root/reference nor this script's inputs are needed.

Contents (all float64):
  inputs    A_K/A_q (set A, n=4096), B_K/B_q (set B, n=4096), N_K/N_q (notebook grid, first 512),
            T_K (64 table energies, 1e-2..1e6 GeV)
  outputs   vmap_<grid>_<element>_<process>           dcs::vmap(f)(...)             4096 / 512 values
            integral_<element>_<process>_<del|cel>_<min_points>   dcs::vmap_integral(recoil_integral)
Elements: rock (STANDARD_ROCK), H, O, Fe, Pb; muon mass.  xlow = X_FRACTION = 0.05.

A second file, tests/golden/coulomb_golden.npz, holds the Coulomb / soft-scattering functions
(dcs::coulomb_data, coulomb_transport, hard_scattering, soft_scattering; dcs.hh:499-952) on
C_K = 256 energies 1e-3..1e6 GeV for the same five elements plus a two-element (water) hard
scattering case:
  cd_<el>_{fcm,screen,fspin,invlambda}, ct_<el>_<mu tag>, hs_<el>_{mu0,lbh}, ss_<el>, C_mu_grid
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from noa_b200 import grids  # noqa: E402

MUON_MASS = 0.10565839
ELEMENTS = {"rock": (22., 0.1364E-6, 11), "H": (1.0087, 19.2E-9, 1), "O": (15.999, 95.0E-9, 8),
            "Fe": (55.845, 286E-9, 26), "Pb": (207.2, 823E-9, 82)}
PROC = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")


def main():
    oracle.build_reference()
    ref = oracle.load_reference()
    assert ref is not None, "oracle/_ref/libnoa_ref.so could not be built (no /root/reference?)"
    out = {}
    out["A_K"], out["A_q"] = grids.set_a(4096)
    out["B_K"], out["B_q"] = grids.set_b(4096)
    K, q = grids.notebook_grid()
    out["N_K"], out["N_q"] = K[:512].copy(), q[:512].copy()
    out["T_K"] = grids.table_energies(64, -2.0, 6.0)
    for g in "ABN":
        for en, el in ELEMENTS.items():
            if g == "N" and en != "rock":
                continue
            for p, pn in enumerate(PROC):
                out[f"vmap_{g}_{en}_{pn}"] = ref.vmap(p, out[g + "_K"], out[g + "_q"], el,
                                                      MUON_MASS, threads=8)
    for en in ("rock", "H", "Pb"):
        for p, pn in enumerate(PROC):
            for ig, ign in enumerate(("del", "cel")):
                for mp in (180, 1000):
                    out[f"integral_{en}_{pn}_{ign}_{mp}"] = ref.vmap_integral(
                        p, ig, out["T_K"], 0.05, mp, ELEMENTS[en], MUON_MASS, threads=8)
    path = os.path.join(ROOT, "tests", "golden", "dcs_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
    coulomb(ref)


MU_CASES = {"one": np.array([1.0]), "small": np.array([1e-3]), "tiny": np.array([1e-12])}


def coulomb(ref):
    out = {}
    K = grids.table_energies(256, -3.0, 6.0)
    out["C_K"] = K
    out["C_mu_grid"] = 10.0 ** np.linspace(-14, 0, K.size)
    data = {}
    for en, el in ELEMENTS.items():
        fcm, scr, fspin, invl = ref.coulomb_data(K, el, MUON_MASS)
        data[en] = (fcm, scr, fspin, invl)
        out[f"cd_{en}_fcm"], out[f"cd_{en}_screen"] = fcm, scr
        out[f"cd_{en}_fspin"], out[f"cd_{en}_invlambda"] = fspin, invl
        for tag, mu in list(MU_CASES.items()) + [("grid", out["C_mu_grid"])]:
            out[f"ct_{en}_{tag}"] = ref.coulomb_transport(scr, fspin, mu)
        mu0, lbh = ref.hard_scattering(out[f"ct_{en}_one"], fcm, scr, invl, fspin)
        out[f"hs_{en}_mu0"], out[f"hs_{en}_lbh"] = mu0, lbh
        out[f"ss_{en}"] = ref.soft_scattering(K, el, MUON_MASS)
    # water: two elements, inverse Wentzel paths weighted by mass fraction
    w = np.array([0.111894, 0.888106])[:, None]
    st = lambda i: np.stack((data["H"][i], data["O"][i]))
    G = np.stack((out["ct_H_one"], out["ct_O_one"]))
    out["hs_water_invlambda"] = st(3) * w
    mu0, lbh = ref.hard_scattering(G, st(0), st(1), out["hs_water_invlambda"], st(2))
    out["hs_water_mu0"], out["hs_water_lbh"] = mu0, lbh
    path = os.path.join(ROOT, "tests", "golden", "coulomb_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
