"""Builds open_whisper_kit_b200/lib/libwhisper.so (C++ host + sm_100a CUDA) in-tree with nvcc.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libwhisper.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
          "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-DWHISPER_SHARED", "-DWHISPER_BUILD",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    out = []
    for dp, _, fns in os.walk(CSRC):
        for fn in sorted(fns):
            if fn.endswith((".cu", ".cpp")):
                out.append(os.path.join(dp, fn))
    return sorted(out)


def _headers_digest():
    h = hashlib.sha1()
    for base in (CSRC, os.path.join(ROOT, "include")):
        for dp, _, fns in os.walk(base):
            for fn in sorted(fns):
                if fn.endswith((".h", ".cuh", ".hpp")):
                    with open(os.path.join(dp, fn), "rb") as f:
                        h.update(fn.encode())
                        h.update(f.read())
    h.update(" ".join(COMMON + ARCH).encode())
    return h.hexdigest()


def _compile(src, hdig, verbose):
    rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
    obj = os.path.join(OBJ, rel + ".o")
    stamp = obj + ".stamp"
    with open(src, "rb") as f:
        dig = hashlib.sha1(f.read() + hdig.encode()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [NVCC] + ARCH + COMMON + (["-x", "cu"] if src.endswith(".cu") else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, r.stderr if verbose else ""


def build(verbose=False, jobs=8):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = _sources()
    hdig = _headers_digest()
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        res = list(ex.map(lambda s: _compile(s, hdig, verbose), srcs))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            if log:
                sys.stderr.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
