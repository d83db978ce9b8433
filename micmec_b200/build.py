"""Build libmicmec_b200.so in-tree with nvcc for sm_100a (B200).  No other architecture is compiled."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmicmec_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=true", "-Xptxas", "-v", "--threads", "0",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    stamp = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "micmec_b200.h"))
    return any(os.path.getmtime(p) > stamp for p in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [nvcc] + flags + ["-o", LIB] + sources()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libmicmec_b200.so")
    with open(os.path.join(HERE, "build.log"), "w") as handle:
        handle.write(" ".join(cmd) + "\n" + proc.stdout)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
