#!/usr/bin/env python
"""Per-operator timing of every reference-shaped module at the C2 shape (256 x 22050), CUDA-graph replayed,
rotating input buffers > L2.  GPU box only.  Prints one JSON line per operator."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_sound_b200.interface.hifi_gan import MelSpectrogram
from pytorch_sound_b200.models import transforms as T

B, L, NBUF = 256, 22050, 8
xs = [torch.randn(B, L, device="cuda") * 0.1 for _ in range(NBUF)]
geo = dict(sample_rate=22050, mel_size=80, n_fft=1024, win_length=1024, hop_length=256, mel_min=0.0, mel_max=8000.0)
ops = {
    "LogMelSpectrogram.forward (clamped)": (T.LogMelSpectrogram(min_db=-50, max_db=30, **geo).cuda(), lambda m, x: m(x)),
    "LogMelSpectrogram.forward(norm=True)": (T.LogMelSpectrogram(min_db=-50, max_db=30, **geo).cuda(), lambda m, x: m(x, norm=True)),
    "hifi_gan.MelSpectrogram.forward": (MelSpectrogram().cuda(), lambda m, x: m(x)),
    "Audio2Mel.forward": (T.Audio2Mel().cuda(), lambda m, x: m(x.unsqueeze(1))),
    "STFT.magnitude (|X| only)": (T.STFT(1024, 256).cuda(), lambda m, x: m.magnitude(x)),
    "STFT.transform (mag, phase)": (T.STFT(1024, 256).cuda(), lambda m, x: m.transform(x)),
    "STFTTorchAudio.forward (re, im)": (T.STFTTorchAudio(1024, 256).cuda(), lambda m, x: m(x)),
}
from pytorch_sound_b200 import functional as _F
ops["MFCC.forward (DCT fused as an epilogue of the mel launch)"] = (
    T.MFCC(22050, 80, 1024, 1024, 40, 256, -50, 30, 0.0, 8000.0).cuda(), lambda m, x: m(x))
ops["MFCC as two launches (mel kernel + DCT kernel)"] = (
    T.MFCC(22050, 80, 1024, 1024, 40, 256, -50, 30, 0.0, 8000.0).cuda(), lambda m, x: _F.mel_to_mfcc(m.mel_func(x), m.dct_mat))
_mag = T.STFT(1024, 256).cuda()
_mags = [_mag.magnitude(x) for x in xs]
ops["LogMelScale.forward (tcgen05 mel GEMM on magnitudes in HBM)"] = (
    T.LogMelScale(22050, 80, 1024, -50, 30, 0.0, 8000.0).cuda(), lambda m, x: m(_mags[[id(t) for t in xs].index(id(x))]))
for name, (mod, fn) in ops.items():
    for x in xs:
        fn(mod, x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            outs = [fn(mod, x) for x in xs]
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * NBUF)
    o = outs[0] if isinstance(outs[0], torch.Tensor) else outs[0]
    out_bytes = sum(t.numel() * 4 for t in (outs[0] if isinstance(outs[0], tuple) else (outs[0],)))
    print(json.dumps({"op": name, "us_per_call": round(us, 2), "hours_audio_per_s": round(B * L / 22050 / 3600 / (us * 1e-6), 1),
                      "out_MB": round(out_bytes / 1e6, 2),
                      "hbm_GBs_read_plus_write": round((4 * B * L + out_bytes) / (us * 1e-6) / 1e9, 1)}))
