"""Build libgdm_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The library travels to the GPU box with the repository snapshot (it is git-ignored,
not gpurun-ignored).  `python -m gdm_b200.build` or `__graft_entry__.build()`.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgdm_b200.so")
SOURCES = ["api.cu", "basis.cpp", "generic.cu", "blas1.cu", "cg.cu", "rk.cu", "comm.cu", "kron3d.cu", "kron3d_pers.cu", "massinv.cu", "cut.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
    "-x", "cu",
]
# GDM_BUILD_EXPERIMENTAL=1: also instantiate the alternative tile shapes of the persistent kernel that were measured in
# round 2 (profiles/r2; selectable with GDM_PERS_CFG).  Doubles the build time.
if os.environ.get("GDM_BUILD_EXPERIMENTAL", "0") == "1":
    FLAGS = FLAGS[:-2] + ["-DGDM_FUSED_EXPERIMENTAL"] + FLAGS[-2:]
# GDM_BUILD_WATCHDOG=1: diagnostic library libgdm_b200_wd.so whose persistent kernel bounds and records every wait
# (select it at run time with GDM_B200_LIB=<path>); the product library is not touched.
BUILD_DIR = "build"
if os.environ.get("GDM_BUILD_WATCHDOG", "0") == "1":
    FLAGS = FLAGS[:-2] + ["-DGDM_PERS_WATCHDOG"] + FLAGS[-2:]
    LIB = os.path.join(HERE, "libgdm_b200_wd.so")
    BUILD_DIR = "build_wd"


# GDM_BUILD_DEFINES="-DX -DY" GDM_BUILD_SUFFIX=_x: A/B builds of kernel variants (libgdm_b200_x.so, selected with GDM_B200_LIB)
if os.environ.get("GDM_BUILD_SUFFIX"):
    FLAGS = FLAGS[:-2] + os.environ.get("GDM_BUILD_DEFINES", "").split() + FLAGS[-2:]
    LIB = os.path.join(HERE, "libgdm_b200" + os.environ["GDM_BUILD_SUFFIX"] + ".so")
    BUILD_DIR = "build" + os.environ["GDM_BUILD_SUFFIX"]


def _stamp():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/gdm/cuda/gdm_c_api.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp_file = os.path.join(HERE, BUILD_DIR, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    os.makedirs(os.path.join(HERE, BUILD_DIR), exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, BUILD_DIR, src.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ldl"]
    subprocess.check_call(cmd)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
