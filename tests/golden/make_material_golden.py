#!/usr/bin/env python
"""Generate tests/golden/material_golden.npz: the per-material table assembly (SURVEY.md 8(f) rank
1) computed by oracle/material_oracle.c DRIVING THE COMPILED REFERENCE's scalar DCS and quadrature
(oracle/ref_shim.cc: noa_ref_dcs_scalar / noa_ref_integral_scalar over the unmodified headers in
/root/reference/src).  Run in the build container; the .npz is committed and travels to the GPU box.

Contents (float64 unless noted), for <m> in water (H + O), rock (standard rock), lead_iron (Pb + Fe,
50/50 -- two heavy elements so that every CSf slot is exercised):
  M_K                     97 energies 1e-2 .. 1e6 GeV
  <m>_elem [ne,3,4,n], <m>_cs [4,n], <m>_cel [4,n], <m>_straggling [n], <m>_csf [ne,4,n],
  <m>_cs_total [n], <m>_xt [ne,4,n], <m>_kt, <m>_it (int)
  a second grid that starts ABOVE every threshold except photonuclear's (M2_K, 1e-1 .. 1e3 GeV, 40
  energies, prefix water2_) so that `it` > 1 never hides the search,
  and the generalised recoil integral on M_K for standard rock:
  mode_<process>_<mode>_<tag>   tag = "cut" (x in [1e-6, 0.05]) or "mid" (x in [0.01, 0.5])
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from noa_b200 import grids  # noqa: E402

MUON_MASS = 0.10565839
H, O = (1.0087, 19.2E-9, 1), (15.999, 95.0E-9, 8)
ROCK, FE, PB = (22., 0.1364E-6, 11), (55.845, 286E-9, 26), (207.2, 823E-9, 82)
MATERIALS = {"water": ((H, O), (0.111894, 0.888106)), "rock": ((ROCK,), (1.0,)),
             "lead_iron": ((PB, FE), (0.5, 0.5))}
PROC = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")


def main():
    oracle.build_reference()
    ref = oracle.load_reference()
    assert ref is not None and ref.kind == "reference"
    out = {"M_K": grids.table_energies(97, -2.0, 6.0), "M2_K": grids.table_energies(40, -1.0, 3.0)}
    for name, (elements, fractions) in MATERIALS.items():
        r = oracle.material_assembly(ref, elements, fractions, MUON_MASS, out["M_K"], 0.05, 180,
                                     threads=8)
        assert r["rc"] == 0
        for key in ("elem", "cs", "cel", "straggling", "csf", "cs_total", "xt"):
            out[f"{name}_{key}"] = r[key]
        out[f"{name}_kt"] = np.array([r["kt"]])
        out[f"{name}_it"] = np.array([r["it"]], dtype=np.int32)
    r = oracle.material_assembly(ref, *MATERIALS["water"], MUON_MASS, out["M2_K"], 0.05, 180,
                                 threads=8)
    for key in ("elem", "cs", "cel", "straggling", "csf", "cs_total", "xt"):
        out[f"water2_{key}"] = r[key]
    out["water2_kt"], out["water2_it"] = np.array([r["kt"]]), np.array([r["it"]], dtype=np.int32)
    for p, pn in enumerate(PROC):
        for mode in (0, 1, 2):
            for tag, (lo, hi) in (("cut", (1E-06, 0.05)), ("mid", (0.01, 0.5))):
                out[f"mode_{pn}_{mode}_{tag}"] = oracle.integral_scalar(
                    ref, p, mode, out["M_K"], lo, hi, ROCK, MUON_MASS, 180)
    path = os.path.join(ROOT, "tests", "golden", "material_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
