"""ctypes bindings for the two CPU checkers (test infrastructure, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libdcs_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libnoa_ref.so")

_dp = ctypes.POINTER(ctypes.c_double)
PROCESSES = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


class Checker:
    """Uniform face over the C port and the compiled reference.

    vmap(process, K, q, element, mass, threads)            -> dcs::vmap / dcs::pvmap
    vmap_integral(process, integrand, K, xlow, n, el, m)   -> dcs::vmap_integral(recoil_integral)
    `element` is (A [g/mol], I [GeV], Z); `integrand` 0 = del (dcs*q), 1 = cel (dcs*q*q).
    """

    def __init__(self, lib, kind):
        self.kind = kind  # "port" | "reference"
        self._lib = lib
        pre = "oracle_" if kind == "port" else "noa_ref_"
        self._vmap = getattr(lib, pre + "vmap")
        self._vint = getattr(lib, pre + "vmap_integral")
        self._vmap.restype = ctypes.c_int
        self._vint.restype = ctypes.c_int
        self._vmap.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int64,
                               ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_double]
        self._vint.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int64,
                               ctypes.c_double, ctypes.c_int32, ctypes.c_double, ctypes.c_double,
                               ctypes.c_int32, ctypes.c_double]
        i64, i32, f64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_double
        self._coulomb = {}
        for name, args in (
                ("coulomb_data", [_dp, _dp, _dp, _dp, _dp, i64, f64, f64, i32, f64]),
                ("coulomb_transport", [_dp, _dp, _dp, _dp, i64, i64]),
                ("hard_scattering", [_dp, _dp, _dp, _dp, _dp, _dp, _dp, i32, i32]),
                ("soft_scattering", [_dp, _dp, i64, f64, f64, i32, f64])):
            fn = getattr(lib, pre + name, None)     # absent in an oracle/_ref built before they existed
            if fn is not None:
                fn.restype = ctypes.c_int
                fn.argtypes = args
                self._coulomb[name] = fn
        thr = getattr(lib, "oracle_max_threads" if kind == "port" else "noa_ref_threads")
        thr.restype = ctypes.c_int
        self.max_threads = int(thr())

    def use_all_cores(self):
        """Size the OpenMP team to the cores this process may run on (torchrun exports
        OMP_NUM_THREADS=1, which would make the 'all host threads' baseline single-threaded)."""
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        if self.kind == "reference" and hasattr(self._lib, "noa_ref_set_threads"):
            self._lib.noa_ref_set_threads(ctypes.c_int(n))
        self.max_threads = n
        return n

    def _threads_arg(self, threads):
        # the port takes a thread count, the reference shim a serial/OpenMP switch
        # (OpenMP team size of the reference = OMP_NUM_THREADS / all cores)
        if self.kind == "port":
            return int(threads)
        return 1 if threads > 1 else 0

    def vmap(self, process, K, q, element, mass, threads=1):
        K, q = _f64(K), _f64(q)
        assert K.size == q.size
        out = np.zeros_like(K)
        rc = self._vmap(int(process), self._threads_arg(threads), _p(K), _p(q), _p(out), K.size,
                        float(element[0]), float(element[1]), int(element[2]), float(mass))
        if rc:
            raise ValueError(f"bad process {process}")
        return out

    def vmap_integral(self, process, integrand, K, xlow, min_points, element, mass, threads=1):
        K = _f64(K)
        out = np.zeros_like(K)
        rc = self._vint(int(process), int(integrand), self._threads_arg(threads), _p(K), _p(out),
                        K.size, float(xlow), int(min_points), float(element[0]),
                        float(element[1]), int(element[2]), float(mass))
        if rc:
            raise ValueError(f"bad process {process}")
        return out


    # ---- Coulomb / soft scattering (src/noa/pms/dcs.hh:499-952) -----------------------------------
    def coulomb_data(self, K, element, mass):
        """dcs::coulomb_data -> (fCM [n,2], screening [n,9], fspin [n], invlambda [n])"""
        K = _f64(K)
        n = K.size
        fcm, scr = np.zeros((n, 2)), np.zeros((n, 9))
        fspin, invl = np.zeros(n), np.zeros(n)
        self._coulomb["coulomb_data"](_p(fcm), _p(scr), _p(fspin), _p(invl), _p(K), n,
                                      float(element[0]), float(element[1]), int(element[2]),
                                      float(mass))
        return fcm, scr, fspin, invl

    def coulomb_transport(self, screening, fspin, mu):
        """dcs::coulomb_transport -> coefficients [n,2]; mu has 1 or n entries"""
        scr, fspin, mu = _f64(screening), _f64(fspin), _f64(mu)
        n = fspin.size
        assert scr.size == 9 * n and mu.size in (1, n)
        coef = np.zeros((n, 2))
        self._coulomb["coulomb_transport"](_p(coef), _p(scr), _p(fspin), _p(mu), mu.size, n)
        return coef

    def hard_scattering(self, G, fcm, screening, invlambda, fspin):
        """dcs::hard_scattering on [nel, nkin, ...] arrays -> (mu0 [nkin], lb_h [nkin])"""
        invl = np.ascontiguousarray(np.asarray(invlambda, dtype=np.float64))
        if invl.ndim == 1:
            invl = invl[None]
        nel, nkin = invl.shape
        G, fcm, scr, fspin = _f64(G), _f64(fcm), _f64(screening), _f64(fspin)
        assert G.size == 2 * nel * nkin and fcm.size == 2 * nel * nkin
        assert scr.size == 9 * nel * nkin and fspin.size == nel * nkin
        mu0, lb_h = np.zeros(nkin), np.zeros(nkin)
        self._coulomb["hard_scattering"](_p(mu0), _p(lb_h), _p(G), _p(fcm), _p(scr),
                                         _p(_f64(invl)), _p(fspin), nel, nkin)
        return mu0, lb_h

    def soft_scattering(self, K, element, mass):
        """dcs::soft_scattering -> ms1 [n]"""
        K = _f64(K)
        out = np.zeros_like(K)
        self._coulomb["soft_scattering"](_p(out), _p(K), K.size, float(element[0]),
                                         float(element[1]), int(element[2]), float(mass))
        return out


def material_assembly(checker, elements, fractions, mass, K, cutoff, min_points, threads=1):
    """Per-material table assembly (oracle/material_oracle.c: the PUMAS algorithm,
    pumas.c:8054-8111, 10768-10808, 10816-10881, 10901-10955) driven by `checker`'s DCS: the
    compiled reference's scalar functions (kind "reference") or the C port's.  Returns a dict of
    arrays: elem [ne,3,4,n] (CSn, cel, stg), cs [4,n], cel [4,n], straggling [n], csf [ne,4,n],
    cs_total [n], xt [ne,4,n] and the scalars kt, it, rc."""
    port = ctypes.CDLL(PORT_SO)
    pre = "oracle_" if checker.kind == "port" else "noa_ref_"
    dcs = getattr(checker._lib, pre + "dcs_scalar")
    integral = getattr(checker._lib, pre + "integral_scalar")
    K = _f64(K)
    n, ne = K.size, len(elements)
    A = np.array([e[0] for e in elements], dtype=np.float64)
    I = np.array([e[1] for e in elements], dtype=np.float64)
    Z = np.array([e[2] for e in elements], dtype=np.int32)
    w = np.array(fractions, dtype=np.float64)
    out = {"elem": np.zeros((ne, 3, 4, n)), "cs": np.zeros((4, n)), "cel": np.zeros((4, n)),
           "straggling": np.zeros(n), "csf": np.zeros((ne, 4, n)), "cs_total": np.zeros(n),
           "xt": np.zeros((ne, 4, n))}
    kt, it = ctypes.c_double(0.), ctypes.c_int32(0)
    fn = port.oracle_material_assembly
    fn.restype = ctypes.c_int
    vp = ctypes.c_void_p
    fn.argtypes = [ctypes.c_int32, _dp, _dp, ctypes.POINTER(ctypes.c_int32), _dp, ctypes.c_double,
                   _dp, ctypes.c_int64, ctypes.c_double, ctypes.c_int32, vp, vp, _dp, _dp, _dp,
                   _dp, _dp, _dp, ctypes.POINTER(ctypes.c_double),
                   ctypes.POINTER(ctypes.c_int32), _dp, ctypes.c_int]
    rc = fn(ne, _p(A), _p(I), Z.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _p(w), float(mass),
            _p(K), n, float(cutoff), int(min_points), ctypes.cast(dcs, vp), ctypes.cast(integral, vp),
            _p(out["elem"]), _p(out["cs"]), _p(out["cel"]), _p(out["straggling"]), _p(out["csf"]),
            _p(out["cs_total"]), ctypes.byref(kt), ctypes.byref(it), _p(out["xt"]), int(threads))
    out["kt"], out["it"], out["rc"] = kt.value, it.value, rc
    return out


def integral_scalar(checker, process, mode, K, xlow, xhigh, element, mass, min_points):
    """The generalised recoil integral (modes 0 / 1 / 2, upper bound x_high) for every energy."""
    pre = "oracle_" if checker.kind == "port" else "noa_ref_"
    fn = getattr(checker._lib, pre + "integral_scalar")
    fn.restype = ctypes.c_double
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                   ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_double,
                   ctypes.c_int32]
    return np.array([fn(int(process), int(mode), float(k), float(xlow), float(xhigh),
                        float(element[0]), float(element[1]), int(element[2]), float(mass),
                        int(min_points)) for k in _f64(K)])


def _make(target):
    subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)


def build_port():
    _make("oracle")
    return PORT_SO


def build_reference(reference_root="/root/reference"):
    """Compile the unmodified reference where it lies; returns None when it is not present."""
    if not os.path.isdir(os.path.join(reference_root, "src", "noa")):
        return REF_SO if os.path.exists(REF_SO) else None
    _make("ref")
    return REF_SO


def load_port():
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < max(
            os.path.getmtime(os.path.join(_HERE, f)) for f in ("dcs_oracle.c", "coulomb_oracle.c")):
        build_port()
    return Checker(ctypes.CDLL(PORT_SO), "port")


def load_reference():
    """The compiled reference, or None if oracle/_ref was never built (it cannot be built on a
    box without /root/reference; the prebuilt .so travels there with gpurun)."""
    if not os.path.exists(REF_SO):
        return None
    try:
        return Checker(ctypes.CDLL(REF_SO), "reference")
    except OSError:
        return None
