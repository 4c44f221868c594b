"""Build recipe for libswegl_b200.so (sm_100a only, in-tree so the .so travels with the repo snapshot)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["abi.cu", "geometry.cu", "fragment.cu", "animate.cu", "../host/image_decode.cpp"]   # the last one is host-only C++ (PNG / JPEG decode)
OUT = os.path.join(HERE, "libswegl_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                       # the reference is built without FMA: keep mul and add separate
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-Xptxas", "-v",
    "--shared",
    "-I", os.path.join(HERE, "..", "include"),
]
LIBS = ["-lz"]                              # zlib inflate for PNG (host/image_decode.cpp)


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "swegl_b200.h"),
                                                               os.path.join(HERE, "host", "image_decode.cpp")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, extra_flags, verbose=False):
    """A/B builds for kernel tuning (tools/ab_variants.sh): libswegl_b200_<name>.so with extra -D flags; select it at
    run time with SWEGL_B200_LIB=<path>.  Not part of the product build."""
    out = os.path.join(HERE, f"libswegl_b200_{name}.so")
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + LIBS
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed building variant {name}")
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES] + LIBS
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libswegl_b200.so")
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:
        f.write("".join(ln for ln in res.stderr.splitlines(True) if "Compile time" not in ln))   # (timings would change the file on every build)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
