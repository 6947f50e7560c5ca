#!/usr/bin/env python
"""Kernel time vs batch size (CUDA-graph replayed launches): separates the fixed launch/prologue cost from the
per-round cost of the persistent kernel.  GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_sound_b200.models.transforms import LogMelSpectrogram

L = int(sys.argv[1]) if len(sys.argv) > 1 else 22050
m = LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.).cuda()
for B in (1, 8, 27, 54, 108, 162, 216, 256, 270, 324, 512, 1024, 2048):
    xs = [torch.randn(B, L, device="cuda") * 0.1 for _ in range(max(2, min(8, int(3e8 // (B * L * 4)))))]
    for x in xs:
        m(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            outs = [m(x) for x in xs]
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(xs))
    T = 1 + L // 256
    tasks = B * ((T + 1) // 2)
    print(f"B={B:5d} tasks={tasks:7d} tasks/warp={tasks / (148 * 16):6.2f}  {us:8.2f} us/launch  {B * L / 22050 / 3600 / (us * 1e-6):9.1f} h/s")
