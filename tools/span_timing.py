#!/usr/bin/env python
"""Kernel start/end (globaltimer) of back-to-back graph-replayed launches: separates the launch gap and the CTA
prologue from the work (GPU box only; builds a -DB200MEL_SPAN_TIMING copy of the library into gpurun_out/)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytorch_sound_b200 import _lib, build

out = os.path.join(ROOT, "gpurun_out", "libb200mel_span.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.run([build.find_nvcc()] + build.NVCC_FLAGS + ["-DB200MEL_SPAN_TIMING", "-o", out, "b200mel.cu"], cwd=build.CSRC, check=True)
_lib.LIB_PATH = out
_lib.lib()
h = C.CDLL(out)
h.b200mel_debug_set_buffer.argtypes = [C.c_void_p]
from pytorch_sound_b200.models.transforms import LogMelSpectrogram

m = LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.).cuda()
for B in (1, 54, 256):
    n = 8
    xs = [torch.randn(B, 22050, device="cuda") * 0.1 for _ in range(n)]
    for x in xs:
        m(x)
    dbg = torch.zeros(n, 20, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for i, x in enumerate(xs):
                h.b200mel_debug_set_buffer(dbg[i].data_ptr())
                m(x)
    h.b200mel_debug_set_buffer(None)
    for rep in range(3):
        dbg.zero_()
        dbg[:, 14] = 2**62
        dbg[:, 16] = 2**62
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
    d = dbg.cpu()
    spans = [(d[i, 15] - d[i, 14]).item() / 1e3 for i in range(n)]
    gaps = [(d[i + 1, 14] - d[i, 15]).item() / 1e3 for i in range(n - 1)]
    pro = [(d[i, 17] - d[i, 14]).item() / 1e3 for i in range(n)]
    period = (d[n - 1, 14] - d[0, 14]).item() / 1e3 / (n - 1)
    print(f"B={B}: period {period:.2f} us | kernel span {sum(spans)/n:.2f} us | gap end->next start {sum(gaps)/(n-1):.2f} us | "
          f"prologue (last warp) {sum(pro)/n:.2f} us")
