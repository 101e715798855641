"""Python mirror of `noa::pms::dcs` for CUDA tensors (reference: src/noa/pms/dcs.hh).

Same names, argument order and semantics as the reference so its tests read the same here:

    dcs.vmap(dcs.pair_production)(result, K, q, STANDARD_ROCK, MUON_MASS)        # dcs.hh:35-48
    r = dcs.map(dcs.photonuclear)(K, q, STANDARD_ROCK, MUON_MASS)                # dcs.hh:50-60
    dcs.vmap_integral(dcs.recoil_integral(dcs.bremsstrahlung, dcs.del_integrand))(
        result, K, dcs.X_FRACTION, STANDARD_ROCK, MUON_MASS, 180)                # dcs.hh:89-130
    dcs.cuda.vmap_bremsstrahlung(result, K, q, STANDARD_ROCK, MUON_MASS)         # dcs.hh:1006-1017

`pvmap` / `pmap` (the reference's OpenMP forms, dcs.hh:62-87) are the same GPU call.  Tensors are
validated (CUDA device, float64, contiguous, equal numel) -- the reference assumes all of that
unchecked (src/noa/utils/common.cuh:45-56).  Work is enqueued on torch's current stream of the
tensors' device without synchronising, like the reference's launch.

Additions with no reference counterpart (it has bremsstrahlung only on the GPU and no fused or
material forms): `cuda.vmap_all`, `cuda.map_all`, `cuda.map_material`, `cuda.tables`.

Everything here calls the C ABI (include/noa_dcs_b200.h); nothing is computed in Python or torch.
"""
import ctypes
import os
from typing import NamedTuple

import torch

from . import _lib
from .physics import AtomicElement, Material, X_FRACTION  # noqa: F401  (X_FRACTION re-exported)


class Process(NamedTuple):
    """Token standing for one of the reference's scalar DCS functors."""
    index: int
    name: str

    def __repr__(self):
        return f"dcs.{self.name}"


bremsstrahlung = Process(0, "bremsstrahlung")      # physics.hh:108-153, dcs.hh:134-140
pair_production = Process(1, "pair_production")    # dcs.hh:144-258
photonuclear = Process(2, "photonuclear")          # dcs.hh:362-405
ionisation = Process(3, "ionisation")              # dcs.hh:408-443
PROCESSES = (bremsstrahlung, pair_production, photonuclear, ionisation)


class Integrand(NamedTuple):
    index: int
    name: str


del_integrand = Integrand(0, "del_integrand")      # dcs * q      (dcs.hh:107-109)
cel_integrand = Integrand(1, "cel_integrand")      # dcs * q * q  (dcs.hh:111-113)


# ---- validation ---------------------------------------------------------------------------------
def _check_tensor(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (noa_b200 has no CPU path); got {t.device}")
    if t.dtype != torch.float64:
        raise ValueError(f"{name} must be float64 (the reference computes in double), got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _same(a, b, na, nb):
    if a.numel() != b.numel():
        raise ValueError(f"{na} and {nb} must have the same number of elements "
                         f"({a.numel()} vs {b.numel()})")
    if a.device != b.device:
        raise ValueError(f"{na} and {nb} must be on the same device")


def _element(element):
    A, I, Z = element
    return float(A), float(I), int(Z)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


# ---- element-wise -------------------------------------------------------------------------------
def _vmap_call(process, result, kinetic_energies, recoil_energies, element, mass):
    lib = _lib.require_device()
    _check_tensor(result, "result")
    _check_tensor(kinetic_energies, "kinetic_energies")
    _check_tensor(recoil_energies, "recoil_energies")
    _same(kinetic_energies, recoil_energies, "kinetic_energies", "recoil_energies")
    _same(kinetic_energies, result, "kinetic_energies", "result")
    A, I, Z = _element(element)
    with torch.cuda.device(kinetic_energies.device):
        _lib.check(lib.noa_dcs_vmap_f64(process.index, _ptr(kinetic_energies),
                                        _ptr(recoil_energies), _ptr(result),
                                        kinetic_energies.numel(), A, I, Z, float(mass),
                                        _stream(kinetic_energies.device)))


def vmap(dcs_func):
    """dcs::vmap (dcs.hh:35-48): writes `result` in place."""
    if not isinstance(dcs_func, Process):
        raise TypeError("dcs.vmap expects one of dcs.bremsstrahlung / pair_production / "
                        "photonuclear / ionisation")

    def apply(result, kinetic_energies, recoil_energies, element, mass):
        _vmap_call(dcs_func, result, kinetic_energies, recoil_energies, element, mass)

    return apply


def map(dcs_func):  # noqa: A001  (the reference's name)
    """dcs::map (dcs.hh:50-60): allocates the result like `kinetic_energies`."""
    run = vmap(dcs_func)

    def apply(kinetic_energies, recoil_energies, element, mass):
        result = torch.empty_like(kinetic_energies)
        run(result, kinetic_energies, recoil_energies, element, mass)
        return result

    return apply


pvmap = vmap   # dcs.hh:62-75
pmap = map     # dcs.hh:77-87


# ---- recoil integrals ---------------------------------------------------------------------------
class RecoilIntegral(NamedTuple):
    """What dcs::recoil_integral(f, integrand) returns in the reference (dcs.hh:89-105, 955-1001):
    a closure over one energy.  Here it is a token consumed by `vmap_integral`."""
    process: Process
    integrand: Integrand


def recoil_integral(dcs_func, integrand):
    if not isinstance(dcs_func, Process) or not isinstance(integrand, Integrand):
        raise TypeError("dcs.recoil_integral(f, integrand): f is a dcs process, integrand is "
                        "dcs.del_integrand or dcs.cel_integrand")
    return RecoilIntegral(dcs_func, integrand)


# Workspaces of the flat table form (16 B per node and process: 641 MB for 10^4 energies x 1002
# nodes), one per (device, stream), grown on demand and kept: taking 641 MB from torch's caching
# allocator on every call cost 0.13 ms of a 4.3 ms build.  Builds queued on one stream run one
# after the other, so they can share it; release_table_workspaces() hands the memory back.
_table_workspaces = {}


def _table_workspace(device, doubles):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _table_workspaces.get(key)
    if ws is None or ws.numel() < doubles:
        ws = torch.empty(doubles, dtype=torch.float64, device=device)
        _table_workspaces[key] = ws
    return ws


def release_table_workspaces():
    """Frees the cached workspaces of cuda.tables() (synchronise first if builds are in flight)."""
    _table_workspaces.clear()


def _table_call(mask, kinetic_energies, xlow, element, mass, min_points, del_out, cel_out,
                flat=True):
    lib = _lib.require_device()
    A, I, Z = _element(element)
    n = kinetic_energies.numel()
    with torch.cuda.device(kinetic_energies.device):
        if flat:
            need = int(lib.noa_dcs_table_workspace_doubles(n, int(min_points)))
            ws = _table_workspace(kinetic_energies.device, need)
            _lib.check(lib.noa_dcs_table_ws_f64(mask, _ptr(kinetic_energies), n, float(xlow),
                                                int(min_points), A, I, Z, float(mass),
                                                _ptr(del_out), _ptr(cel_out), _ptr(ws), need,
                                                _stream(kinetic_energies.device)))
            return
        _lib.check(lib.noa_dcs_table_f64(mask, _ptr(kinetic_energies), n,
                                         float(xlow), int(min_points), A, I, Z, float(mass),
                                         _ptr(del_out) if del_out is not None else None,
                                         _ptr(cel_out) if cel_out is not None else None,
                                         _stream(kinetic_energies.device)))


def vmap_integral(cs_integral):
    """dcs::vmap_integral (dcs.hh:115-130): result[i] = cs_integral(K[i], xlow, element, mass,
    min_points)."""
    if not isinstance(cs_integral, RecoilIntegral):
        raise TypeError("dcs.vmap_integral expects dcs.recoil_integral(f, integrand)")

    def apply(result, kinetic_energies, xlow, element, mass, min_points):
        _check_tensor(result, "result")
        _check_tensor(kinetic_energies, "kinetic_energies")
        _same(kinetic_energies, result, "kinetic_energies", "result")
        lib = _lib.require_device()
        A, I, Z = _element(element)
        with torch.cuda.device(kinetic_energies.device):
            _lib.check(lib.noa_dcs_vmap_integral_f64(
                cs_integral.process.index, cs_integral.integrand.index, _ptr(kinetic_energies),
                _ptr(result), kinetic_energies.numel(), float(xlow), int(min_points), A, I, Z,
                float(mass), _stream(kinetic_energies.device)))

    return apply


# ---- Coulomb scattering and soft scattering (dcs.hh:499-952) ------------------------------------
NSF = 9    # physics.hh:86


def _numel(t, name, expected):
    _check_tensor(t, name)
    if t.numel() != expected:
        raise ValueError(f"{name} must hold {expected} elements, has {t.numel()}")


def coulomb_data(fCM, screening, fspin, invlambda, kinetic_energies, element, mass):
    """dcs::coulomb_data (dcs.hh:600-622): fills fCM [n, 2], screening [n, 9], fspin [n],
    invlambda [n] for the energies [n]."""
    lib = _lib.require_device()
    _check_tensor(kinetic_energies, "kinetic_energies")
    n = kinetic_energies.numel()
    _numel(fCM, "fCM", 2 * n)
    _numel(screening, "screening", NSF * n)
    _numel(fspin, "fspin", n)
    _numel(invlambda, "invlambda", n)
    A, I, Z = _element(element)
    with torch.cuda.device(kinetic_energies.device):
        _lib.check(lib.noa_dcs_coulomb_data_f64(_ptr(kinetic_energies), n, A, I, Z, float(mass),
                                                _ptr(fCM), _ptr(screening), _ptr(fspin),
                                                _ptr(invlambda), _stream(kinetic_energies.device)))


def coulomb_transport(coefficients, screening, fspin, mu):
    """dcs::coulomb_transport (dcs.hh:674-693): coefficients [n, 2]; `mu` holds one cutoff for all
    energies or one per energy."""
    lib = _lib.require_device()
    _check_tensor(fspin, "fspin")
    n = fspin.numel()
    _numel(coefficients, "coefficients", 2 * n)
    _numel(screening, "screening", NSF * n)
    _check_tensor(mu, "mu")
    if mu.numel() not in (1, n):
        raise ValueError(f"mu must hold 1 or {n} elements, has {mu.numel()}")
    with torch.cuda.device(fspin.device):
        _lib.check(lib.noa_dcs_coulomb_transport_f64(_ptr(screening), _ptr(fspin), _ptr(mu),
                                                     mu.numel(), n, _ptr(coefficients),
                                                     _stream(fspin.device)))


def hard_scattering(mu0, lb_h, coefficients, transform, screening, invlambdas, fspins):
    """dcs::hard_scattering (dcs.hh:843-872): invlambdas / fspins [nel, nkin], coefficients /
    transform [nel, nkin, 2], screening [nel, nkin, 9]; writes mu0 [nkin] and lb_h [nkin]."""
    lib = _lib.require_device()
    _check_tensor(invlambdas, "invlambdas")
    if invlambdas.dim() != 2:
        raise ValueError("invlambdas must be [nel, nkin]")
    nel, nkin = invlambdas.shape
    _numel(fspins, "fspins", nel * nkin)
    _numel(coefficients, "coefficients", 2 * nel * nkin)
    _numel(transform, "transform", 2 * nel * nkin)
    _numel(screening, "screening", NSF * nel * nkin)
    _numel(mu0, "mu0", nkin)
    _numel(lb_h, "lb_h", nkin)
    with torch.cuda.device(invlambdas.device):
        _lib.check(lib.noa_dcs_hard_scattering_f64(_ptr(coefficients), _ptr(transform),
                                                   _ptr(screening), _ptr(invlambdas), _ptr(fspins),
                                                   int(nel), int(nkin), _ptr(mu0), _ptr(lb_h),
                                                   _stream(invlambdas.device)))


def soft_scattering(ms1, kinetic_energies, element, mass):
    """dcs::soft_scattering (dcs.hh:940-952): ms1[i] for every energy."""
    lib = _lib.require_device()
    _check_tensor(kinetic_energies, "kinetic_energies")
    n = kinetic_energies.numel()
    _numel(ms1, "ms1", n)
    A, I, Z = _element(element)
    with torch.cuda.device(kinetic_energies.device):
        _lib.check(lib.noa_dcs_soft_scattering_f64(_ptr(kinetic_energies), n, A, I, Z, float(mass),
                                                   _ptr(ms1), _stream(kinetic_energies.device)))


# ---- noa::pms::dcs::cuda ------------------------------------------------------------------------
class _Cuda:
    """`noa::pms::dcs::cuda` (dcs.hh:1004-1019, src/noa/pms/dcs.cuh:30-51).  The reference has
    bremsstrahlung only; the other three processes and the fused forms are new."""

    # -- reference surface
    @staticmethod
    def vmap_bremsstrahlung(result, kinetic_energies, recoil_energies, element, mass):
        _vmap_call(bremsstrahlung, result, kinetic_energies, recoil_energies, element, mass)

    @staticmethod
    def map_bremsstrahlung(kinetic_energies, recoil_energies, element, mass):
        # dcs.cuh:43-51 (zeros_like + vmap; every element is overwritten, so empty_like here)
        return map(bremsstrahlung)(kinetic_energies, recoil_energies, element, mass)

    # -- same shape for the other processes
    @staticmethod
    def vmap_pair_production(result, kinetic_energies, recoil_energies, element, mass):
        _vmap_call(pair_production, result, kinetic_energies, recoil_energies, element, mass)

    @staticmethod
    def map_pair_production(kinetic_energies, recoil_energies, element, mass):
        return map(pair_production)(kinetic_energies, recoil_energies, element, mass)

    @staticmethod
    def vmap_photonuclear(result, kinetic_energies, recoil_energies, element, mass):
        _vmap_call(photonuclear, result, kinetic_energies, recoil_energies, element, mass)

    @staticmethod
    def map_photonuclear(kinetic_energies, recoil_energies, element, mass):
        return map(photonuclear)(kinetic_energies, recoil_energies, element, mass)

    @staticmethod
    def vmap_ionisation(result, kinetic_energies, recoil_energies, element, mass):
        _vmap_call(ionisation, result, kinetic_energies, recoil_energies, element, mass)

    @staticmethod
    def map_ionisation(kinetic_energies, recoil_energies, element, mass):
        return map(ionisation)(kinetic_energies, recoil_energies, element, mass)

    # -- fused: the four processes in one pass, result[p] = process p
    @staticmethod
    def vmap_all(result, kinetic_energies, recoil_energies, element, mass):
        lib = _lib.require_device()
        _check_tensor(result, "result")
        _check_tensor(kinetic_energies, "kinetic_energies")
        _check_tensor(recoil_energies, "recoil_energies")
        _same(kinetic_energies, recoil_energies, "kinetic_energies", "recoil_energies")
        n = kinetic_energies.numel()
        if result.numel() != 4 * n:
            raise ValueError(f"result must hold 4 x {n} elements, has {result.numel()}")
        A, I, Z = _element(element)
        with torch.cuda.device(kinetic_energies.device):
            _lib.check(lib.noa_dcs_vmap_all_f64(_ptr(kinetic_energies), _ptr(recoil_energies),
                                                _ptr(result), n, A, I, Z, float(mass),
                                                _stream(kinetic_energies.device)))

    @staticmethod
    def map_all(kinetic_energies, recoil_energies, element, mass):
        result = torch.empty((4,) + tuple(kinetic_energies.shape), dtype=torch.float64,
                             device=kinetic_energies.device)
        _Cuda.vmap_all(result, kinetic_energies, recoil_energies, element, mass)
        return result

    # -- material = mass-fraction mix of elements (water = H + O)
    @staticmethod
    def vmap_material(result, kinetic_energies, recoil_energies, material, mass,
                      processes=PROCESSES):
        lib = _lib.require_device()
        _check_tensor(result, "result")
        _check_tensor(kinetic_energies, "kinetic_energies")
        _check_tensor(recoil_energies, "recoil_energies")
        _same(kinetic_energies, recoil_energies, "kinetic_energies", "recoil_energies")
        n = kinetic_energies.numel()
        mask = 0
        for pr in processes:
            mask |= 1 << pr.index
        nproc = bin(mask).count("1")
        if result.numel() != nproc * n:
            raise ValueError(f"result must hold {nproc} x {n} elements, has {result.numel()}")
        ne = len(material.elements)
        A = (ctypes.c_double * ne)(*[float(e.A) for e in material.elements])
        I = (ctypes.c_double * ne)(*[float(e.I) for e in material.elements])
        Z = (ctypes.c_int32 * ne)(*[int(e.Z) for e in material.elements])
        w = (ctypes.c_double * ne)(*[float(f) for f in material.fractions])
        with torch.cuda.device(kinetic_energies.device):
            _lib.check(lib.noa_dcs_vmap_mixture_f64(mask, _ptr(kinetic_energies),
                                                    _ptr(recoil_energies), _ptr(result), n, ne, A,
                                                    I, Z, w, float(mass),
                                                    _stream(kinetic_energies.device)))

    @staticmethod
    def map_material(kinetic_energies, recoil_energies, material, mass, processes=PROCESSES):
        nproc = len({pr.index for pr in processes})
        result = torch.empty((nproc,) + tuple(kinetic_energies.shape), dtype=torch.float64,
                             device=kinetic_energies.device)
        _Cuda.vmap_material(result, kinetic_energies, recoil_energies, material, mass, processes)
        return result

    # -- fused table builder: DEL and CEL of all requested processes from one DCS evaluation/node
    @staticmethod
    def tables(kinetic_energies, xlow, element, mass, min_points, processes=PROCESSES,
               out=None, flat=None):
        """Returns (del, cel), each [4, n_K]; rows of processes not requested are zero.
        `flat`: True = the workspace form (noa_dcs_table_ws_f64), False = one CTA per row
        (noa_dcs_table_f64), None = True unless NOA_DCS_TABLE_FLAT=0; same bits either way."""
        if flat is None:
            flat = os.environ.get("NOA_DCS_TABLE_FLAT", "1") != "0"
        _check_tensor(kinetic_energies, "kinetic_energies")
        n = kinetic_energies.numel()
        mask = 0
        for pr in processes:
            mask |= 1 << pr.index
        if out is None:
            del_t = torch.zeros((4, n), dtype=torch.float64, device=kinetic_energies.device)
            cel_t = torch.zeros((4, n), dtype=torch.float64, device=kinetic_energies.device)
        else:
            del_t, cel_t = out
            _check_tensor(del_t, "del")
            _check_tensor(cel_t, "cel")
            if del_t.numel() != 4 * n or cel_t.numel() != 4 * n:
                raise ValueError("out tensors must hold 4 x n_K elements each")
        if n:
            _table_call(mask, kinetic_energies, xlow, element, mass, min_points, del_t, cel_t,
                        flat=flat)
        return del_t, cel_t


    # -- material tables: element tables mixed by mass fraction (pumas.c:8054-8078)
    @staticmethod
    def material_tables(kinetic_energies, xlow, material, mass, min_points, processes=PROCESSES,
                        scratch=None):
        """Returns ([2, 4, n_K] material table, [n_elements, 2, 4, n_K] element tables)."""
        lib = _lib.require_device()
        _check_tensor(kinetic_energies, "kinetic_energies")
        n = kinetic_energies.numel()
        ne = len(material.elements)
        mask = 0
        for pr in processes:
            mask |= 1 << pr.index
        dev = kinetic_energies.device
        if scratch is None:
            scratch = torch.empty((ne, 2, 4, n), dtype=torch.float64, device=dev)
        else:
            _check_tensor(scratch, "scratch")
            if scratch.numel() != ne * 8 * n:
                raise ValueError(f"scratch must hold {ne} x 8 x {n} elements")
        table = torch.zeros((2, 4, n), dtype=torch.float64, device=dev)
        A = (ctypes.c_double * ne)(*[float(e.A) for e in material.elements])
        I = (ctypes.c_double * ne)(*[float(e.I) for e in material.elements])
        Z = (ctypes.c_int32 * ne)(*[int(e.Z) for e in material.elements])
        w = (ctypes.c_double * ne)(*[float(f) for f in material.fractions])
        if n:
            with torch.cuda.device(dev):
                need = int(lib.noa_dcs_table_workspace_doubles(n, int(min_points)))
                ws = _table_workspace(dev, need)
                _lib.check(lib.noa_dcs_table_material_f64(
                    mask, _ptr(kinetic_energies), n, float(xlow), int(min_points), ne, A, I, Z, w,
                    float(mass), _ptr(scratch), _ptr(table), _ptr(ws), need, _stream(dev)))
        return table, scratch.view(ne, 2, 4, n)


    # -- PUMAS-style table assembly of a material over NOA's DCS
    @staticmethod
    def vmap_integral_mode(result, kinetic_energies, process, mode, xlow, xhigh, element, mass,
                           min_points):
        """The recoil integral in PUMAS's compute_dcs_integral shape (pumas.c:10901-10955):
        integrand dcs q^(1 + mode) in ln q over [ln(K xlow), ln(K xhigh)], / (K + mass).
        mode 0 cross-section, 1 energy loss, 2 straggling."""
        lib = _lib.require_device()
        _check_tensor(result, "result")
        _check_tensor(kinetic_energies, "kinetic_energies")
        _same(kinetic_energies, result, "kinetic_energies", "result")
        A, I, Z = _element(element)
        with torch.cuda.device(kinetic_energies.device):
            _lib.check(lib.noa_dcs_vmap_integral_mode_f64(
                process.index, int(mode), _ptr(kinetic_energies), _ptr(result),
                kinetic_energies.numel(), float(xlow), float(xhigh), int(min_points), A, I, Z,
                float(mass), _stream(kinetic_energies.device)))

    @staticmethod
    def material_assembly(kinetic_energies, cutoff, material, mass, min_points=180):
        """pumas.c:8054-8111, 10768-10808, 10816-10881 over NOA's DCS, all on the GPU.  Returns a
        dict of device tensors: elem [ne, 3, 4, n] (CSn, cel, stg per element), cs / cel [4, n],
        straggling [n], csf [ne, 4, n], cs_total [n], kt [1], it [1] (int32), xt [ne, 4, n]."""
        lib = _lib.require_device()
        _check_tensor(kinetic_energies, "kinetic_energies")
        n = kinetic_energies.numel()
        ne = len(material.elements)
        dev = kinetic_energies.device
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)  # noqa: E731
        out = {"elem": z(ne, 3, 4, n), "cs": z(4, n), "cel": z(4, n), "straggling": z(n),
               "csf": z(ne, 4, n), "cs_total": z(n), "kt": z(1),
               "it": torch.zeros(1, dtype=torch.int32, device=dev), "xt": z(ne, 4, n)}
        A = (ctypes.c_double * ne)(*[float(e.A) for e in material.elements])
        I = (ctypes.c_double * ne)(*[float(e.I) for e in material.elements])
        Z = (ctypes.c_int32 * ne)(*[int(e.Z) for e in material.elements])
        w = (ctypes.c_double * ne)(*[float(f) for f in material.fractions])
        if n:
            with torch.cuda.device(dev):
                need = int(lib.noa_dcs_table_workspace_doubles(n, int(min_points)))
                ws = _table_workspace(dev, need)
                _lib.check(lib.noa_dcs_material_assembly_f64(
                    _ptr(kinetic_energies), n, float(cutoff), int(min_points), ne, A, I, Z, w,
                    float(mass), _ptr(out["elem"]), _ptr(out["cs"]), _ptr(out["cel"]),
                    _ptr(out["straggling"]), _ptr(out["csf"]), _ptr(out["cs_total"]),
                    _ptr(out["kt"]), _ptr(out["it"]), _ptr(out["xt"]), _ptr(ws), need,
                    _stream(dev)))
        return out


cuda = _Cuda()


# ---- host-buffer path (the CPU-tensor drop-in: dcs::map(f) on CPU tensors) ------------------------
class HostStager:
    """Runs CPU tensors through the GPU.  Pinned tensors (one process): a single kernel reads and
    writes the host arrays in place over PCIe (C ABI: noa_dcs_vmap_pinned_f64).  Pageable tensors,
    or all four processes at once: chunked H2D / kernel / D2H pipeline over the stager's device
    scratch and streams (C ABI: noa_dcs_vmap_host_f64).  Both block until `out` is complete."""

    def __init__(self, chunk_pairs=1 << 18, n_slots=3, device=None, zero_copy=True):
        self._lib = _lib.require_device()
        self.zero_copy = zero_copy
        self._handle = ctypes.c_void_p()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.noa_dcs_stager_create(ctypes.byref(self._handle),
                                                      int(chunk_pairs), int(n_slots)))

    def close(self):
        if self._handle:
            self._lib.noa_dcs_stager_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def map(self, dcs_func, kinetic_energies, recoil_energies, element, mass, out=None):  # noqa: A003
        """CPU float64 tensors in, CPU tensor out (blocking).  dcs_func=None evaluates all four
        processes (out is [4, n])."""
        for t, name in ((kinetic_energies, "kinetic_energies"), (recoil_energies, "recoil_energies")):
            if t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous float64 CPU tensor")
        n = kinetic_energies.numel()
        if recoil_energies.numel() != n:
            raise ValueError("kinetic_energies and recoil_energies differ in size")
        index = 4 if dcs_func is None else dcs_func.index
        shape = ((4,) if dcs_func is None else ()) + tuple(kinetic_energies.shape)
        if out is None:
            out = torch.empty(shape, dtype=torch.float64,
                              pin_memory=kinetic_energies.is_pinned())
        else:
            # the C ABI writes n (or 4 n) doubles through the raw pointer: refuse anything else
            expected = (4 if dcs_func is None else 1) * n
            if (not isinstance(out, torch.Tensor) or out.is_cuda or out.dtype != torch.float64
                    or not out.is_contiguous() or out.numel() != expected):
                raise ValueError(f"out must be a contiguous float64 CPU tensor of {expected} "
                                 "elements")
        A, I, Z = _element(element)
        if (dcs_func is not None and self.zero_copy and kinetic_energies.is_pinned()
                and recoil_energies.is_pinned() and out.is_pinned() and out.is_contiguous()):
            # page-locked buffers: the kernel reads and writes them in place over PCIe, no copies
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream(self.device)
                _lib.check(self._lib.noa_dcs_vmap_pinned_f64(
                    index, _ptr(kinetic_energies), _ptr(recoil_energies), _ptr(out), n, A, I, Z,
                    float(mass), ctypes.c_void_p(stream.cuda_stream)))
                stream.synchronize()
            return out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.noa_dcs_vmap_host_f64(self._handle, index, _ptr(kinetic_energies),
                                                      _ptr(recoil_energies), _ptr(out), n, A, I, Z,
                                                      float(mass)))
        return out
