"""Build libkoifish_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m koifish_b200.build [--force] [--verbose]

Sources: csrc/Device/*.cu (kernels + C ABI), csrc/Tensor/*.cpp, csrc/Transformer/*.cpp (host runtime), csrc/TokenSet/*.cpp (tokenizer, host only).
"""
import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
TAG = os.environ.get("KF_BUILD_TAG", "")  # tuning builds: KF_BUILD_TAG=dbg KF_NVCC_DEFS=-DKF_DEBUG_KNOBS -> libkoifish_b200_dbg.so (load with KF_LIB_PATH)
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libkoifish_b200%s.so" % ("_" + TAG if TAG else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CSRC, "Device"),
    "-diag-suppress", "177",
] + os.environ.get("KF_NVCC_DEFS", "").split()  # e.g. KF_NVCC_DEFS=-DKF_GEMV_OCC=4 for tuning experiments
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _sources():
    cu = sorted(glob.glob(os.path.join(CSRC, "Device", "*.cu")))
    cpp = sorted(glob.glob(os.path.join(CSRC, "Tensor", "*.cpp")) + glob.glob(os.path.join(CSRC, "Transformer", "*.cpp")) +
                 glob.glob(os.path.join(CSRC, "TokenSet", "*.cpp")))
    return cu, cpp


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    cu, cpp = _sources()
    headers = glob.glob(os.path.join(CSRC, "**", "*.h*"), recursive=True) + glob.glob(os.path.join(CSRC, "**", "*.cuh"), recursive=True) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest(cu + cpp + headers)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = [NVCC, "-ccbin", HOST_CXX] + NVCC_FLAGS + ["-x", "cu" if src.endswith(".cu") else "c++", "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, cu + cpp))
    cmd = [NVCC, "-ccbin", HOST_CXX, "-shared", "--cudart", "static", "-o", LIB] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
