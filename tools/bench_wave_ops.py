#!/usr/bin/env python
"""Timing of the waveform-side / MFCC kernels against their HBM roofline (GPU box only).

    python tools/bench_wave_ops.py

C2-shaped inputs (256 x 22050 samples; mel 256 x 80 x 87), rotating over enough buffers to exceed L2, CUDA-graph
replayed, CUDA events.  Prints one JSON line per operator with the algorithmic bytes and the achieved GB/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from pytorch_sound_b200.models.sound import PreEmphasis
from pytorch_sound_b200.models.transforms import MelToMFCC
from pytorch_sound_b200.utils.calculate import volume_norm_log_torch

peak = 6531.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, inputs, reps=40):
    for x in inputs:
        fn(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            outs = [fn(x) for x in inputs]
    torch.cuda.synchronize()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(inputs)) * 1e3, outs


B, L, M, T, C = 256, 22050, 80, 87, 40
for scale in (1, 16):  # C2 batch and a 16x larger one (launch-latency vs streaming regime)
    n = max(2, 12 // scale)
    wavs = [torch.randn(B * scale, 1, L, device="cuda") * 0.1 for _ in range(n)]
    mels = [torch.randn(B * scale, M, T, device="cuda") for _ in range(n)]
    pe, mf = PreEmphasis().cuda(), MelToMFCC(C, M).cuda()
    rows = [("PreEmphasis.forward", lambda x: pe(x), wavs, 8 * B * scale * L),
            ("volume_norm_log_torch", lambda x: volume_norm_log_torch(x), wavs, 12 * B * scale * L),
            ("MelToMFCC.forward", lambda x: mf(x), mels, 4 * B * scale * T * (M + C))]
    for name, fn, ins, nbytes in rows:
        us, _ = timed(fn, ins)
        print(json.dumps({"op": name, "clips": B * scale, "us_per_call": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 2),
                          "GBs": round(nbytes / us / 1e3, 1), "frac_of_measured_hbm": round(nbytes / us / 1e3 / peak, 3)}))
