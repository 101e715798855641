#!/usr/bin/env python
"""Generate tests/golden/special_golden.npz from the COMPILED REFERENCE (oracle/_ref): the four DCS
and their DEL/CEL integrals on inputs outside anything physical -- zeros, signed zero, subnormals,
the kinematic thresholds themselves, q >= K, huge values, infinities, NaN, negative energies.  The
reference has no input validation: whatever its arithmetic yields there (0, NaN, inf,
garbage) is the contract, and it is where a kernel that folds its special-case handling is most
likely to differ.  Run in the build container; the .npz is committed.

  S_K, S_q                          all pairs of the special values below
  vmap_S_<element>_<process>        dcs::vmap(f)(...)
  ST_K                              special table energies
  integral_S_<element>_<process>_<del|cel>_180
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

MUON_MASS = 0.10565839
ELEMENTS = {"rock": (22., 0.1364E-6, 11), "H": (1.0087, 19.2E-9, 1), "Pb": (207.2, 823E-9, 82)}
PROC = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")
ME = 0.510998910E-03
SPECIAL = np.array([
    0.0, -0.0, 5e-324, 1e-310, 2.2250738585072014e-308, 1e-300, 1e-100, 1e-30, 1e-10, 0.62 * 0.1364E-6,
    1e-3, 4. * ME, np.nextafter(4. * ME, 1.0), 0.05, MUON_MASS, 0.134977 * (1. + 0.134977 / (2 * 0.931494)),
    1.0, 2.0, 10.818, 1e3, 1e3 + MUON_MASS, 1e8, 1e15, 1e100, 1e300, 1.7e308, np.inf, np.nan, -1.0, -1e3])


def main():
    oracle.build_reference()
    ref = oracle.load_reference()
    assert ref is not None, "oracle/_ref/libnoa_ref.so could not be built (no /root/reference?)"
    out = {}
    K, q = np.meshgrid(SPECIAL, SPECIAL, indexing="ij")
    out["S_K"], out["S_q"] = K.reshape(-1).copy(), q.reshape(-1).copy()
    out["ST_K"] = np.array([0.0, 5e-324, 1e-300, 1e-10, 1e-3, 0.05, MUON_MASS, 1.0, 10.818, 10.819,
                            1e3, 1e8, 1e15, 1e100, 1e300, np.inf, np.nan, -1.0])
    with np.errstate(all="ignore"):
        for en, el in ELEMENTS.items():
            for p, pn in enumerate(PROC):
                out[f"vmap_S_{en}_{pn}"] = ref.vmap(p, out["S_K"], out["S_q"], el, MUON_MASS, threads=1)
                for ig, ign in enumerate(("del", "cel")):
                    out[f"integral_S_{en}_{pn}_{ign}_180"] = ref.vmap_integral(
                        p, ig, out["ST_K"], 0.05, 180, el, MUON_MASS, threads=1)
    path = os.path.join(ROOT, "tests", "golden", "special_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
    for pn in PROC:
        v = out[f"vmap_S_rock_{pn}"]
        print(pn, "nan", int(np.isnan(v).sum()), "inf", int(np.isinf(v).sum()), "zero", int((v == 0).sum()),
              "finite nonzero", int((np.isfinite(v) & (v != 0)).sum()))


if __name__ == "__main__":
    main()
