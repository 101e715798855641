#!/usr/bin/env python
"""Build the LibTorch boundary in-tree with the host compiler (no nvcc needed):
  noa_b200/libnoa_dcs_b200_torch.so   noa::pms::dcs::cuda::* for C++ callers (torch_api.cc)
  noa_b200/_muons.so                  pybind11 module mirroring docs/pms/muon_dcs.{cc,cu}
  noa_b200/measure_dcs_calc_cuda      the reference's benchmark cases (benchmark/measure-dcs-calc*.cc)
Both link libnoa_dcs_b200.so with an $ORIGIN rpath, so the tree is relocatable (gpurun copies it).
"""
import os
import subprocess
import sys
import sysconfig

import torch
from torch.utils import cpp_extension

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def newer(target, sources):
    return os.path.exists(target) and all(os.path.getmtime(s) <= os.path.getmtime(target)
                                          for s in sources)


def main(force=False):
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = ["-I" + p for p in cpp_extension.include_paths()]
    cuda_inc = "/usr/local/cuda/include"
    if os.path.isdir(cuda_inc):
        inc.append("-I" + cuda_inc)
    inc.append("-I" + sysconfig.get_paths()["include"])
    common = [CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=1",
              "-DTORCH_API_INCLUDE_EXTENSION_H"] + inc
    link = ["-L" + tlib, "-L" + PKG, "-lnoa_dcs_b200", "-ltorch", "-ltorch_cpu", "-lc10",
            "-lc10_cuda", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib]
    hdr = [os.path.join(PKG, "..", "include", "noa_b200", "pms_dcs_cuda.hh"),
           os.path.join(PKG, "..", "include", "noa_b200", "pms_dcs.hh"),
           os.path.join(PKG, "..", "include", "noa_dcs_b200.h")]

    api_src = os.path.join(HERE, "torch_api.cc")
    api_out = os.path.join(PKG, "libnoa_dcs_b200_torch.so")
    if force or not newer(api_out, [api_src] + hdr):
        # pure C++ / LibTorch: must not pull in the Python bindings
        plain = [a for a in common if a != "-DTORCH_API_INCLUDE_EXTENSION_H"]
        subprocess.run(plain + [api_src, "-o", api_out] + link, check=True)

    ext_src = os.path.join(HERE, "muon_dcs_ext.cc")
    ext_out = os.path.join(PKG, "_muons.so")
    if force or not newer(ext_out, [ext_src, api_out] + hdr):
        subprocess.run(common + ["-DTORCH_EXTENSION_NAME=_muons", ext_src, "-o", ext_out] + link +
                       ["-lnoa_dcs_b200_torch", "-ltorch_python"], check=True)
    # the reference's benchmark cases as an executable (benchmark/measure_dcs_calc_cuda.cc)
    bench_src = os.path.join(PKG, "..", "benchmark", "measure_dcs_calc_cuda.cc")
    bench_out = os.path.join(PKG, "measure_dcs_calc_cuda")
    if os.path.exists(bench_src) and (force or not newer(bench_out, [bench_src, api_out] + hdr)):
        exe = [a for a in common if a not in ("-shared", "-DTORCH_API_INCLUDE_EXTENSION_H")]
        tmp_out = bench_out + ".tmp"       # a failed link must not leave a stale target behind
        subprocess.run(exe + [bench_src, "-o", tmp_out] + link +
                       ["-lnoa_dcs_b200_torch", "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart"],
                       check=True)
        os.replace(tmp_out, bench_out)
    print("built", api_out, "and", ext_out)


if __name__ == "__main__":
    main(force="--force" in sys.argv)
