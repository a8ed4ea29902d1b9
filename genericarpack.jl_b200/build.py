"""In-tree build of libgarpack_b200.so (hand-written CUDA for sm_100a + the C ABI).

    python genericarpack.jl_b200/build.py [--force]

Every .cu under csrc/ is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
(in parallel) and linked into genericarpack.jl_b200/libgarpack_b200.so.  The .so is git-ignored
but travels to the GPU box with the gpurun snapshot.  No torch / no JIT cache involved.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libgarpack_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# development A/B builds: GA_B200_DEFS="-DGA_DOT2_FAST" GA_B200_VARIANT=alt -> libgarpack_b200_alt.so (loaded with GA_B200_LIB=...)
EXTRA = os.environ.get("GA_B200_DEFS", "").split()
_VARIANT = os.environ.get("GA_B200_VARIANT", "")
if _VARIANT:
    OBJ = os.path.join(HERE, "build_" + _VARIANT)
    SO = os.path.join(HERE, "libgarpack_b200_%s.so" % _VARIANT)
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
    "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr",
]


def _deps():
    hdr = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "host", "*.hpp"))
    hdr.append(os.path.join(HERE, "..", "include", "garpack_b200.h"))
    return hdr


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [NVCC] + FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, obj, p.returncode, p.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    newest_hdr = max(os.path.getmtime(h) for h in _deps())
    todo = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or verbose or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), newest_hdr):
            todo.append(s)
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, obj, rc, out in ex.map(lambda s: _compile(s, verbose), todo):
                if out.strip() and (rc != 0 or verbose):
                    sys.stderr.write("== %s\n%s\n" % (os.path.basename(src), out))
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s" % src)
    if todo or not os.path.exists(SO):
        cmd = [NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
                                                    "-Xcompiler", "-fPIC", "-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
