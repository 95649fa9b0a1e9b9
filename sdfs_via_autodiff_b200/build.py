"""Build libsdfs_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python sdfs_via_autodiff_b200/build.py [--force] [--verbose]   (runs without importing the package)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsdfs_b200.so")
SOURCES = ["api.cu", "builder.cu", "ops.cu", "kron_apply.cu", "loops.cu", "loops_dense.cu", "loops_kron_sa.cu", "loops_kron_newton.cu",
           "loops_kron_anderson.cu", "loops_cont.cu",
           "comm.cu", "sweep.cu", "small.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    if any(not os.path.exists(os.path.join(CSRC, src.replace(".cu", ".o"))) for src in SOURCES):
        return True                          # an object is missing: the library was not produced by this script's last run
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".o")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "sdfs_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "sdfs_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    for src in SOURCES:                       # one nvcc process per translation unit, all in parallel
        path = os.path.join(CSRC, src)
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_time):
            continue                          # object newer than its source and every header
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
