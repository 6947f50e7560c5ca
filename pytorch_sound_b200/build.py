"""In-tree nvcc build of libb200mel.so (sm_100a only).

`python -m pytorch_sound_b200.build` or `__graft_entry__.build()`.  The .so lands next to this
file so it travels with the source tree; nothing is installed into site-packages.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libb200mel.so")
SOURCES = ["b200mel.cu"]
HEADERS = ["fft32.cuh", "logmel_kernel.cuh", "logmel_fast.cuh", "mel_tc.cuh", "stft_tc.cuh", "stft_tc_tables.h", "spec_kernel.cuh", "wave_ops.cuh", os.path.join("..", "..", "include", "b200mel.h")]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libb200mel.so cannot be built")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    import fcntl

    # several ranks of one torchrun job may get here at once: one compiles, the others wait and reuse the result
    with open(os.path.join(PKG_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB_PATH
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
            res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)  # atomic: a concurrently loading process never sees a half-written file
            if verbose:
                sys.stderr.write(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
