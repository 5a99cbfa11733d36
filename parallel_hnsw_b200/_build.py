"""In-tree build of the C-ABI library (libphnsw.so) with nvcc for sm_100a.

`python -m parallel_hnsw_b200._build` or `__graft_entry__.build()`.  The built .so sits next
to this file so that it travels with the repo snapshot to the GPU box.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# developer knobs: PHNSW_LIB_OUT = path of a variant library (selected at run time with
# PHNSW_LIB), PHNSW_EXTRA_FLAGS = extra nvcc flags (e.g. -DPHNSW_TREE_WARPS=24) for it
LIB_PATH = os.environ.get("PHNSW_LIB_OUT") or os.path.join(_HERE, "libphnsw.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-shared", "-cudart", "static",
] + os.environ.get("PHNSW_EXTRA_FLAGS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _deps():
    out = _sources()
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(_HERE), "include", "phnsw.h"))
    return out


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False):
    """Compile every .cu/.cpp under csrc/ into libphnsw.so (one object per source, in parallel)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(_HERE, "build", os.path.basename(LIB_PATH).replace(".so", "")
                          if os.environ.get("PHNSW_LIB_OUT") else "")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-dc" if False else "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", tmp] + objs
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
