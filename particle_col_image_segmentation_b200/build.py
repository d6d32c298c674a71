"""Build libpcs.so (sm_100a) in-tree with nvcc.  No JIT, no torch arch list."""

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpcs.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
    "-I", os.path.join(HERE, "..", "include"),
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc")) + [os.path.join(HERE, "..", "include", "pcs.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, extra_flags=()):
    """``variant`` / ``extra_flags``: an A/B build with extra ``-D`` flags into ``libpcs_<variant>.so`` (load it with
    ``PCS_LIB_PATH``); tuning only, the product is the plain ``libpcs.so``."""
    lib = LIB if not variant else os.path.join(HERE, f"libpcs_{variant}.so")
    if not variant and not force and not needs_build():
        return LIB
    objs = []
    bdir = os.path.join(HERE, "build" if not variant else f"build_{variant}")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, *extra_flags, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libpcs.so")
    subprocess.check_call([NVCC, "-shared", "-o", lib, *objs, "-lcudart"])
    return lib


if __name__ == "__main__":
    variant = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant, extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
