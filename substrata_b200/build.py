"""In-tree build of libsubstrata_b200.so (the C-ABI product library: host engine + sm_100a kernels).

nvcc cross-compiles for sm_100a without a GPU; the built .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  `python -m substrata_b200.build` or `__graft_entry__.build()` run this.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libsubstrata_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"] + ARCH
# (source, extra flags).  k1_params.cu: exact IEEE rounding, results feed truncations (see the file header).
CUDA_UNITS = [
    ("k1_params.cu", ["-fmad=false"]),
    ("k2_mix.cu", []),
    ("device_cuda.cu", []),
    ("k5_resampler.cu", []),
    ("k6_polyphase.cu", []),
]
# Host units: no FMA contraction -- sh_table.cc must evaluate the encoder table like the reference's SSE build.
HOST_UNITS = ["engine.cc", "sh_table.cc", "polyphase.cc"]
HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-Wall", "-Wextra", "-Wno-unused-parameter"]


def find_nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, defines=(), out=None):
    """`defines` / `out`: development variants (e.g. ("-DSBM_STAGES=4",) -> another .so for A/B runs on the GPU box)."""
    global OBJ, LIB
    if out:
        OBJ, LIB = out + "_obj", out
        force = True
    nvcc = find_nvcc()
    cuda_home = os.path.dirname(os.path.dirname(nvcc))
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "sbm_audio.h"))
    objs = []
    log = []
    for src, extra in CUDA_UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [nvcc] + NVCC_COMMON + list(defines) + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed:\n" + log[-1])
    for src in HOST_UNITS:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = ["g++"] + HOST_FLAGS + ["-I", os.path.join(cuda_home, "include"), "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("g++ failed:\n" + log[-1])
    if force or _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("link failed:\n" + log[-1])
    host_lib = build_host_library(objs_dir=OBJ, lib=LIB, force=force, log=log)
    with open(os.path.join(OBJ, "build.log"), "a") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


HOST_DIR = os.path.join(HERE, "host")


def build_host_library(objs_dir, lib, force=False, log=None):
    """libsubstrata_b200_host.so: the glare::AudioEngine / glare::AudioResampler mirror (host/*.cpp) above the C ABI."""
    out = os.path.join(os.path.dirname(lib), "libsubstrata_b200_host.so")
    srcs = [os.path.join(HOST_DIR, f) for f in ("AudioEngine.cpp", "AudioResampler.cpp", "StreamerThread.cpp")]
    deps = srcs + [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR) if f.endswith(".h")] + [lib]
    if force or _newer(out, deps):
        cmd = ["g++"] + HOST_FLAGS + ["-fvisibility=default", "-shared", "-o", out] + srcs + [
            "-L" + os.path.dirname(lib), "-l:" + os.path.basename(lib), "-Wl,-rpath,$ORIGIN", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if log is not None:
            log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("host library build failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
