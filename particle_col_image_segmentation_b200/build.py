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


def _deps():
    return sorted(sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc")) + [os.path.join(HERE, "..", "include", "pcs.h")])


def source_hash():
    """Hash of every file the library is built from.  Stored next to the library by ``build()`` and compared by
    ``_lib.load()``: a library older than its sources is rebuilt (or refused), never silently used.  Content, not
    mtimes: a snapshot copied to another box keeps the former and not necessarily the latter."""
    import hashlib

    h = hashlib.sha256()
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(f for f in FLAGS if not os.path.isabs(f) and not f.startswith(HERE)).encode())  # flags, not paths: the tree moves between boxes
    return h.hexdigest()


def built_hash():
    try:
        with open(LIB + ".hash") as f:
            return f.read().strip()
    except OSError:
        return None


def needs_build():
    return not os.path.exists(LIB) or built_hash() != source_hash()


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
    # the link step gets the arch too: without it nvcc adds an (empty) device-link stub for its default sm_52
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcudart"])
    if not variant:
        with open(LIB + ".hash", "w") as f:
            f.write(source_hash() + "\n")
    return lib


if __name__ == "__main__":
    variant = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant, extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
