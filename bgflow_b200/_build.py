"""In-tree nvcc build of ``libbgflow_b200.so`` for sm_100a (no JIT cache, no torch headers).

Every ``csrc/*.cu`` is compiled to an object file under ``csrc/_obj/`` (in parallel, only when the
source or any header is newer than the object) and the objects are linked into one shared library.
"""

import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.path.join(_HERE, "libbgflow_b200.so")

# BGX_PAIR_INSTRUMENT=1 in the environment compiles the pair kernel's timing experiments / wait accounting in
# (tools/trace_pair.py); the shipped library is built without them
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
] + (["-DBGX_PAIR_INSTRUMENT=1"] if os.environ.get("BGX_PAIR_INSTRUMENT") == "1" else [])


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libbgflow_b200.so")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return res.stderr


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` and link one shared library.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    newest_header = max(os.path.getmtime(h) for h in _headers())
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (force or not os.path.exists(obj)
                or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)):
            cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for log in pool.map(_run, jobs):
            if verbose:
                print(log)
    for stale in set(glob.glob(os.path.join(OBJ, "*.o"))) - set(objs):
        os.remove(stale)
    _run([nvcc, "-shared", "-o", LIB_PATH, *objs])
    return LIB_PATH
