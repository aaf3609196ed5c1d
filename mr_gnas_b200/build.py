"""Builds libmrgnas.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m mr_gnas_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmrgnas.so")
SOURCES = ["rowops.cu", "segreduce.cu", "graph.cu", "gemm_tc.cu", "amax_bwd.cu", "gate_pipe.cu", "score.cu", "mixed_pre.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
if os.environ.get("MRG_TC_PROF"):       # MMA-warp wait counters in the tcgen05 kernel (scripts/prof_amax_tc.py)
    NVCC_FLAGS.append("-DMRG_TC_PROF")
if os.environ.get("MRG_DW_PROF"):       # phase counters in the a_max dW kernel (scripts/prof_amax_bwd.py)
    NVCC_FLAGS.append("-DMRG_DW_PROF")


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha256()
    files = _sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    files.append(os.path.join(INCLUDE, "mrgnas.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "libmrgnas.digest")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
