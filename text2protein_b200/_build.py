"""Builds libt2p.so (the C-ABI CUDA library, include/t2p.h) in-tree with nvcc for sm_100a.

Used by ``__graft_entry__.build()``; nvcc cross-compiles without a GPU.  The built library sits next to this
file so that it travels with the repo snapshot to the GPU box.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libt2p.so")
OBJ_DIR = os.path.join(os.path.dirname(HERE), "build", "obj")

SOURCES = ["api.cu", "unet.cu", "gemm_tc.cu", "gemm_simt.cu", "norm.cu", "attention.cu", "attention_mma.cu", "attention_tc.cu",
           "elementwise.cu", "pc_step.cu", "conditions.cu", "final_conv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--use_fast_math=false"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


LIB_KNOBS = os.path.join(HERE, "libt2p_knobs.so")


def build(force=False, verbose=False, knobs=False):
    """``knobs=True`` builds libt2p_knobs.so with -DT2P_TIMING_KNOBS: the A/B and timing switches of tools/ are
    environment variables that only that build reads (csrc/common.h env_knob)."""
    if knobs:
        return _build(force, verbose, LIB_KNOBS, os.path.join(os.path.dirname(HERE), "build", "obj_knobs"),
                      ["-DT2P_TIMING_KNOBS"])
    return _build(force, verbose, LIB, OBJ_DIR, [])


def _build(force, verbose, LIB, OBJ_DIR, defines):
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "t2p.h"))
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    stamp = os.path.join(OBJ_DIR, "stamp.txt")
    want = _digest(headers + [os.path.join(CSRC, s) for s in sources])
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        key = _digest(headers + [os.path.join(CSRC, src)])
        keyfile = obj + ".key"
        if not force and os.path.exists(obj) and os.path.exists(keyfile) and open(keyfile).read() == key:
            return obj
        flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
        extra = (["-DT2P_HAVE_ATTENTION_MMA"] if "attention_mma.cu" in sources else []) + defines
        cmd = [_nvcc()] + flags + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        with open(keyfile, "w") as f:
            f.write(key)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


def build_selftest():
    """Stand-alone device self-test of the implicit-GEMM kernels (build/selftest_gemm)."""
    out = os.path.join(os.path.dirname(HERE), "build", "selftest_gemm")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [_nvcc()] + flags + ["-o", out] + [os.path.join(CSRC, s) for s in ("selftest_gemm.cu", "gemm_tc.cu", "gemm_simt.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"selftest build failed:\n{r.stdout}\n{r.stderr}")
    return out


def build_probe():
    """Hardware probe of row-shifted swizzled UMMA operands (build/probe_shift), see csrc/probe_shift.cu."""
    out = os.path.join(os.path.dirname(HERE), "build", "probe_shift")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [_nvcc()] + flags + ["-o", out, os.path.join(CSRC, "probe_shift.cu"), "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"probe build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force=False, verbose=True))
