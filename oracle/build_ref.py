#!/usr/bin/env python
"""Build oracle/_ref: the reference's OWN native coder, compiled from where it lies.

TEST INFRASTRUCTURE ONLY. Nothing under cra5_b200/ imports or links this.

Compiles, unmodified and in place (no sources are copied into this repo):
  /root/reference/cra5/models/compressai/cpp_exts/rans/rans_interface.cpp -> _ref/compressai/ans.<ext>.so
  /root/reference/cra5/models/compressai/cpp_exts/ops/ops.cpp             -> _ref/compressai/_CXX.<ext>.so
with the flags of the reference's (disabled) setup.py:71-75 (`-std=c++17 -O3`).
The only missing piece, ryg_rans `rans64.h` (un-vendored, setup.py:68), is supplied by
oracle/rans64_restated.h installed as _ref/include/rans64.h.

Outputs go only into oracle/_ref/ (git-ignored, NOT gpurun-ignored: the .so files
travel to the GPU box, /root/reference does not).
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CRA5_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def have_reference():
    return os.path.isfile(os.path.join(REF, "cra5/models/compressai/cpp_exts/rans/rans_interface.cpp"))


def build(verbose=True):
    if not have_reference():
        if verbose:
            print(f"[build_ref] {REF} not present; keeping prebuilt oracle/_ref as is")
        return False
    import pybind11

    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = os.path.join(OUT, "include")
    pkg = os.path.join(OUT, "compressai")
    os.makedirs(inc, exist_ok=True)
    os.makedirs(pkg, exist_ok=True)
    shutil.copyfile(os.path.join(HERE, "rans64_restated.h"), os.path.join(inc, "rans64.h"))
    cpp = os.path.join(REF, "cra5/models/compressai/cpp_exts")
    common = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
              f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}"]
    jobs = [
        (["-I" + inc, "-I" + os.path.join(cpp, "rans"), os.path.join(cpp, "rans/rans_interface.cpp")],
         os.path.join(pkg, "ans" + ext)),
        ([os.path.join(cpp, "ops/ops.cpp")], os.path.join(pkg, "_CXX" + ext)),
    ]
    for args, out in jobs:
        src = args[-1]
        if os.path.exists(out) and os.path.getmtime(out) >= max(
                os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "rans64_restated.h"))):
            continue
        cmd = common + args + ["-o", out]
        if verbose:
            print("[build_ref]", " ".join(cmd))
        subprocess.check_call(cmd)
    return True


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok or os.path.isdir(OUT) else 1)
