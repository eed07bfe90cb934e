"""Build libosiris_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m osiris_b200.build [--force]

nvcc cross-compiles without a GPU.  The library links cudart statically and loads libnccl with
dlopen, so it has no link-time dependency beyond libdl/librt/pthread and loads on a box without a
GPU (where only the topology helpers are usable).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libosiris_b200.so")
SOURCES = ["push.cu", "fields.cu", "particles.cu", "comm.cu", "sim.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-fmad=false",  # every fused multiply-add is explicit: parity with the reference statement
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: build libosiris_b200_<variant>.so with extra -D flags (A/B runs of kernel
    variants, selected at run time with OSGPU_LIB); the default build takes neither."""
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "osiris_b200.h"))
    objdir = os.path.join(HERE, "build" if not variant else f"build_{variant}")
    out = OUT if not variant else os.path.join(HERE, f"libosiris_b200_{variant}.so")
    flags = FLAGS + [f"-D{d}" for d in defines]
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([NVCC] + flags + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(out, objs):
        run([NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                   "-Xcompiler", "-fPIC", "-ldl", "-lpthread", "-lrt"])
    return out


if __name__ == "__main__":
    # python -m osiris_b200.build [--force] [--variant NAME -DX=1 -DY=2 ...]
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose=True, variant=var,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
