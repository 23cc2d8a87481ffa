"""Build liblina_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liblina_b200.so")
DEBUG_CSRC = os.path.join(CSRC, "debug")
DEBUG_LIB = os.path.join(LIBDIR, "liblina_b200_debug.so")     # bring-up probes (include/lina_b200_debug.h), not the product
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build liblina_b200.so")
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "lina_b200.h"))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "lina_b200_debug.h"))
    objs, procs = [], []
    for s in sources():
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    # the bring-up probes compile into their own shared object (with their own copy of the error-string helper)
    # ... together with a copy of the GLA chunk kernel compiled with its clock64 timeline (-DLINA_GLA_TRACE) and every other
    # product object (the traced kernel calls into the recurrence fallback and reads the variant table)
    traced = os.path.join(objdir, "debug_gla_chunk_sm100_traced.o")
    dbg_objs = [o for o in objs if not o.endswith(os.sep + "gla_chunk_sm100.o")] + [traced]
    tsrc = os.path.join(CSRC, "gla_chunk_sm100.cu")
    if force or _stale(traced, [tsrc] + headers):
        cmd = [nvcc] + NVCC_FLAGS + ["-DLINA_GLA_TRACE", "-c", tsrc, "-o", traced]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append(("gla_chunk_sm100.cu (traced)", subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s in sorted(f for f in os.listdir(DEBUG_CSRC) if f.endswith(".cu")):
        src, obj = os.path.join(DEBUG_CSRC, s), os.path.join(objdir, "debug_" + s[:-3] + ".o")
        dbg_objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    relink = bool(procs)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    for lib, lobjs in ((LIB, objs), (DEBUG_LIB, dbg_objs)):
        if force or relink or _stale(lib, lobjs):
            cmd = [nvcc, "-shared", "-o", lib] + lobjs + ["-gencode", "arch=compute_100a,code=sm_100a"]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
