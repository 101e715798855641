"""ctypes binding of the C ABI declared in include/noa_dcs_b200.h.

The shared library is built in-tree (noa_b200/libnoa_dcs_b200.so, see noa_b200/csrc/Makefile).
Loading fails loudly if it is missing, or if the host libm is not the one the kernels restate
(noa_dcs_selfcheck; NOA_DCS_ALLOW_LIBM_MISMATCH=1 turns that into a warning); calling fails loudly if
there is no CUDA device.  There is no fallback of any kind.

The measurement kernels (FP64 pipe probes, alternative lane mappings) are a separate library,
noa_b200/libnoa_dcs_b200_probe.so (include/noa_dcs_b200_probe.h): `load_probe()`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NOA_DCS_LIB: developer hook to load an alternative build of the same library (kernel experiments)
LIB_PATH = os.environ.get("NOA_DCS_LIB") or os.path.join(_HERE, "libnoa_dcs_b200.so")

_c_double_p = ctypes.c_void_p   # raw addresses (device or host)
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_f64 = ctypes.c_double
_vp = ctypes.c_void_p

# name -> (restype, argtypes); must list every function include/noa_dcs_b200.h declares
SIGNATURES = {
    "noa_dcs_abi_version": (ctypes.c_int, []),
    "noa_dcs_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "noa_dcs_device_count": (ctypes.c_int, []),
    "noa_dcs_vmap_f64": (ctypes.c_int, [ctypes.c_int, _vp, _vp, _vp, _i64, _f64, _f64, _i32, _f64,
                                        _vp]),
    "noa_dcs_vmap_all_f64": (ctypes.c_int, [_vp, _vp, _vp, _i64, _f64, _f64, _i32, _f64, _vp]),
    "noa_dcs_vmap_mixture_f64": (ctypes.c_int, [ctypes.c_uint, _vp, _vp, _vp, _i64, _i32,
                                                ctypes.POINTER(_f64), ctypes.POINTER(_f64),
                                                ctypes.POINTER(_i32), ctypes.POINTER(_f64), _f64,
                                                _vp]),
    "noa_dcs_table_f64": (ctypes.c_int, [ctypes.c_uint, _vp, _i64, _f64, _i32, _f64, _f64, _i32,
                                         _f64, _vp, _vp, _vp]),
    "noa_dcs_table_material_f64": (ctypes.c_int, [ctypes.c_uint, _vp, _i64, _f64, _i32, _i32,
                                                  ctypes.POINTER(_f64), ctypes.POINTER(_f64),
                                                  ctypes.POINTER(_i32), ctypes.POINTER(_f64), _f64,
                                                  _vp, _vp, _vp, _i64, _vp]),
    "noa_dcs_table_scatter_f64": (ctypes.c_int, [ctypes.c_uint, _vp, _i64, _f64, _i32, _f64, _f64,
                                                 _i32, _f64, _i32, ctypes.POINTER(_vp),
                                                 ctypes.POINTER(_vp), _i64, _i64, _i64, _vp]),
    "noa_dcs_table_exchange_f64": (ctypes.c_int, [ctypes.c_uint, _vp, _i64, _f64, _i32, _f64, _f64,
                                                  _i32, _f64, _i32, _i32, ctypes.POINTER(_vp),
                                                  ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp,
                                                  _vp, _vp, _vp, _vp, _i64, ctypes.c_uint32, _i64,
                                                  _i64, _i64, _f64, _vp]),
    "noa_dcs_table_workspace_doubles": (_i64, [_i64, _i32]),
    "noa_dcs_table_ws_f64": (ctypes.c_int, [ctypes.c_uint, _vp, _i64, _f64, _i32, _f64, _f64, _i32,
                                            _f64, _vp, _vp, _vp, _i64, _vp]),
    "noa_dcs_allgather_f64": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "noa_dcs_vmap_integral_f64": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _vp, _i64, _f64,
                                                 _i32, _f64, _f64, _i32, _f64, _vp]),
    "noa_dcs_vmap_integral_mode_f64": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _vp, _i64,
                                                      _f64, _f64, _i32, _f64, _f64, _i32, _f64,
                                                      _vp]),
    "noa_dcs_material_assembly_f64": (ctypes.c_int, [_vp, _i64, _f64, _i32, _i32,
                                                     ctypes.POINTER(_f64), ctypes.POINTER(_f64),
                                                     ctypes.POINTER(_i32), ctypes.POINTER(_f64),
                                                     _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                     _vp, _vp, _i64, _vp]),
    "noa_dcs_coulomb_data_f64": (ctypes.c_int, [_vp, _i64, _f64, _f64, _i32, _f64, _vp, _vp, _vp,
                                                _vp, _vp]),
    "noa_dcs_coulomb_transport_f64": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "noa_dcs_hard_scattering_f64": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp,
                                                   _vp]),
    "noa_dcs_soft_scattering_f64": (ctypes.c_int, [_vp, _i64, _f64, _f64, _i32, _f64, _vp, _vp]),
    "noa_dcs_stager_create": (ctypes.c_int, [ctypes.POINTER(_vp), _i64, _i32]),
    "noa_dcs_stager_destroy": (ctypes.c_int, [_vp]),
    "noa_dcs_vmap_host_f64": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _i64, _f64, _f64,
                                             _i32, _f64]),
    "noa_dcs_vmap_pinned_f64": (ctypes.c_int, [ctypes.c_int, _vp, _vp, _vp, _i64, _f64, _f64, _i32,
                                               _f64, _vp]),
    "noa_dcs_selfcheck": (ctypes.c_int, [ctypes.POINTER(_i64)]),
    "noa_dcs_launch_info": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_i32),
                                           ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
    "noa_dcs_launch_count": (_i64, []),
    "noa_dcs_div_recomputes": (ctypes.c_int, [ctypes.POINTER(_i64), ctypes.c_int]),
}


# include/noa_dcs_b200_probe.h
PROBE_LIB_PATH = os.environ.get("NOA_DCS_PROBE_LIB") or os.path.join(_HERE,
                                                                    "libnoa_dcs_b200_probe.so")
PROBE_SIGNATURES = {
    "noa_dcs_fp64_probe": (ctypes.c_int, [_i64, _i32, _i32, _vp, _vp]),
    "noa_dcs_fp64_probe_mode": (ctypes.c_int, [_i32, _i64, _i32, _i32, _vp, _vp]),
    "noa_dcs_probe_pair_lanes_f64": (ctypes.c_int, [_vp, _vp, _vp, _i64, _f64, _f64, _i32, _f64,
                                                    _vp]),
}


class NoaDcsError(RuntimeError):
    pass


_lib = None
_probe = None


def load():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NoaDcsError(
            f"{LIB_PATH} is missing: build it with `make -C noa_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "noa_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.noa_dcs_abi_version() != 3:
        raise NoaDcsError("libnoa_dcs_b200.so ABI version mismatch")
    bad = _i64(0)
    if lib.noa_dcs_selfcheck(ctypes.byref(bad)) != 0:
        msg = (f"host libm differs from the glibc (>= 2.28, FMA variant) the kernels restate "
               f"({bad.value} mismatches): results would no longer be bit-identical to the "
               "reference's CPU path")
        if os.environ.get("NOA_DCS_ALLOW_LIBM_MISMATCH") == "1":
            import warnings
            warnings.warn(msg)
        else:
            raise NoaDcsError(msg + " (set NOA_DCS_ALLOW_LIBM_MISMATCH=1 to continue anyway)")
    _lib = lib
    return lib


def load_probe():
    """The measurement library (bench.py's FP64 peak probe, tools/)."""
    global _probe
    if _probe is not None:
        return _probe
    if not os.path.exists(PROBE_LIB_PATH):
        raise NoaDcsError(f"{PROBE_LIB_PATH} is missing: build it with `make -C noa_b200/csrc`")
    lib = ctypes.CDLL(PROBE_LIB_PATH)
    for name, (res, args) in PROBE_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _probe = lib
    return lib


def check(code):
    if code != 0:
        msg = load().noa_dcs_strerror(code)
        raise NoaDcsError(f"noa_dcs error {code}: {msg.decode() if msg else '?'}")


def require_device():
    lib = load()
    if lib.noa_dcs_device_count() <= 0:
        raise NoaDcsError("no CUDA device visible: noa_b200 runs on a B200 only "
                          "(there is no CPU implementation in this package)")
    return lib
