#!/usr/bin/env python
"""Per-phase cycle accounting of the fused kernel (GPU box only).

Builds a -DB200MEL_PHASE_TIMING copy of the library into gpurun_out/ (the shipped .so is untouched), runs one
launch per batch size and prints the average clock64() cycles per task spent in each phase of a warp."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from pytorch_sound_b200 import _lib, build

out = os.path.join(ROOT, "gpurun_out", "libb200mel_timing.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = [build.find_nvcc()] + build.NVCC_FLAGS + ["-DB200MEL_PHASE_TIMING", "-o", out, "b200mel.cu"]
subprocess.run(cmd, cwd=build.CSRC, check=True)
_lib.LIB_PATH = out
lib = _lib.lib()
handle = C.CDLL(out)
handle.b200mel_debug_set_buffer.argtypes = [C.c_void_p]

from pytorch_sound_b200.models.transforms import LogMelSpectrogram

names = ["prologue", "decode", "wait TMA", "stage->regs", "fft pass 1", "xpose+twiddle", "prefetch issue",
         "fft pass 2", "separation", "mel: round setup", "mel: FMAs", "mel: log epilogue", "mel: stores",
         "mel: final sync"]
m = LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.).cuda()
L = 22050
for B in (1, 27, 54, 256, 2048):
    x = torch.randn(B, L, device="cuda") * 0.1
    m(x)
    torch.cuda.synchronize()
    dbg = torch.zeros(20, dtype=torch.int64, device="cuda")
    dbg[14] = 2**62
    dbg[16] = 2**62
    handle.b200mel_debug_set_buffer(dbg.data_ptr())
    m(x)
    torch.cuda.synchronize()
    handle.b200mel_debug_set_buffer(None)
    d = dbg.cpu().tolist()
    print(f"   globaltimer: kernel span {(d[15] - d[14]) / 1e3:.2f} us; prologue done first warp +{(d[16] - d[14]) / 1e3:.2f} us, "
          f"last warp +{(d[17] - d[14]) / 1e3:.2f} us")
    tasks = B * 44
    warps = min(148 * 16, tasks)
    print(f"B={B}: tasks {tasks}, tasks/warp {tasks / (148 * 16):.2f}")
    print(f"   prologue {d[0] / (148 * 16):9.0f} cycles per warp")
    tot = sum(d[1:14])
    for i in range(1, 14):
        print(f"   {names[i]:16s} {d[i] / tasks:9.0f} cycles/task  {100 * d[i] / tot:5.1f} %")
    print(f"   total            {tot / tasks:9.0f} cycles/task")
