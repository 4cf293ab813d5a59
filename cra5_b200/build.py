"""Build libcra5b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m cra5_b200.build [--force]

The shared library lands in cra5_b200/lib/ (git-ignored; it travels to the GPU box with the snapshot).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libcra5b200.so")
VARIANTS = {"": []}   # extra -D flag sets build libcra5b200_<name>.so next to the default (none at present)


def lib_path(variant=""):
    return os.path.join(OUT_DIR, f"libcra5b200{'_' + variant if variant else ''}.so")


def _obj_dir(variant=""):
    return os.path.join(OUT_DIR, "obj" + ("_" + variant if variant else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(f for f in os.listdir(SRC) if f.endswith((".cu", ".cpp")))


def _digest(path, flags):
    h = hashlib.sha1()
    # any header change rebuilds everything: cheap and safe
    for f in sorted(os.listdir(SRC)):
        if f.endswith((".h", ".cuh")):
            h.update(open(os.path.join(SRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "cra5_b200.h"), "rb").read())
    h.update(open(path, "rb").read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile(src, variant=""):
    flags = NVCC_FLAGS + VARIANTS[variant]
    path = os.path.join(SRC, src)
    obj = os.path.join(_obj_dir(variant), src + ".o")
    stamp = obj + ".sha1"
    dig = _digest(path, flags)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = ["nvcc"] + flags + ["-x", "cu", "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    return obj, True


def build(force=False, verbose=True, variant=""):
    obj_dir, lib = _obj_dir(variant), lib_path(variant)
    os.makedirs(obj_dir, exist_ok=True)
    if force:
        for f in os.listdir(obj_dir):
            os.remove(os.path.join(obj_dir, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s_: _compile(s_, variant), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(c for _, c in results)
    if rebuilt or not os.path.exists(lib):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[cra5_b200.build] built {lib}")
    elif verbose:
        print(f"[cra5_b200.build] up to date: {lib}")
    return lib


def build_all(force=False, verbose=True):
    return [build(force=force, verbose=verbose, variant=v) for v in VARIANTS]


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
