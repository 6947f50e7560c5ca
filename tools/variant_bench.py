#!/usr/bin/env python
"""A/B timing of compile-time kernel variants (GPU box only).

    python tools/variant_bench.py "" "-DB200MEL_WARPS_PER_CTA=12" "ENV:B200MEL_NO_FAST=1" "ENV:B200MEL_TC=1" ...

Every argument is a set of extra nvcc flags; each variant is built into gpurun_out/ (the shipped .so is untouched),
loaded in a fresh subprocess and timed the way bench.py times `value`: a CUDA graph of 8 launches over 8 distinct
C2 batches (256 x 22050), replayed, CUDA events.  Prints us per launch and the largest difference from the default
build's output.  (The upper-bound probes of round 1 — no filterbank, no twiddle loads, ... — were removed from the kernel
after their results went into DESIGN.md section 4.)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, torch
sys.path.insert(0, %(root)r)
from pytorch_sound_b200 import _lib
_lib.LIB_PATH = %(lib)r
from pytorch_sound_b200.models.transforms import LogMelSpectrogram
B, L = %(B)d, %(L)d
m = LogMelSpectrogram(*%(geo)r).cuda()
g0 = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.randn(B, L, device="cuda", generator=g0) * 0.1 for _ in range(8)]
for x in xs: m(x)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        outs = [m(x) for x in xs]
torch.cuda.synchronize()
for _ in range(5): g.replay()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 400 * 1e3)
torch.save(outs[0].cpu(), %(out)r)
print("%%.2f" %% best)
'''


def main():
    from pytorch_sound_b200 import build
    # VB_WORKLOAD=C2 (default) | C4 (n_fft 2048 split-mode kernel) | C5 (16 kHz, full-spectrum pair kernel)
    wl = os.environ.get("VB_WORKLOAD", "C2")
    geo, B0, L0 = {"C2": ((22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.), 256, 22050),
                   "C4": ((44100, 128, 2048, 2048, 512, -50, 30, 0., None), 16, 441000),
                   "C5": ((16000, 80, 1024, 1024, 256, -50, 30, 0., 8000.), 8192, 8000)}[wl]
    B = int(os.environ.get("VB_CLIPS", str(B0)))
    L = int(os.environ.get("VB_L", str(L0)))
    ref = None
    for i, flags in enumerate(sys.argv[1:] or [""]):
        lib = os.path.join(ROOT, "gpurun_out", f"libb200mel_var{i}.so")
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        # words of the form ENV:NAME=VALUE set an environment variable of the timed child instead of an nvcc flag
        env = dict(os.environ)
        words = []
        for w in flags.split():
            if w.startswith("ENV:"):
                k, _, v = w[4:].partition("=")
                env[k] = v
            else:
                words.append(w)
        cmd = [build.find_nvcc()] + build.NVCC_FLAGS + words + ["-o", lib, "b200mel.cu"]
        r = subprocess.run(cmd, cwd=build.CSRC, capture_output=True, text=True)
        if r.returncode:
            print(f"[{flags}] BUILD FAILED\n{r.stderr[-2000:]}")
            continue
        out = os.path.join(ROOT, "gpurun_out", f"var{i}.pt")
        r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, lib=lib, out=out, B=B, L=L, geo=geo)],
                           capture_output=True, text=True, env=env)
        if r.returncode:
            print(f"[{flags}] RUN FAILED\n{r.stderr[-2000:]}")
            continue
        import torch
        y = torch.load(out)
        if ref is None:
            ref = y
        err = (y - ref).abs().max().item()
        print(f"[{flags or 'default'}] {r.stdout.strip()} us/launch   max|y - y_default| = {err:.3g}", flush=True)
        os.remove(lib)


if __name__ == "__main__":
    main()
