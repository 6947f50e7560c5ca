#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel variant once, with edge
frames, ragged lengths and a batch larger than one wave of warps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_sound_b200.interface.hifi_gan import MelSpectrogram
from pytorch_sound_b200.models import transforms as T
from pytorch_sound_b200.models.sound import PreEmphasis
from pytorch_sound_b200.utils.calculate import volume_norm_log_torch

torch.manual_seed(0)
x = torch.randn(70, 6001, device="cuda") * 0.1
lens = torch.randint(600, 6001, (70,), dtype=torch.int32, device="cuda")
lm = T.LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.).cuda()
outs = [lm(x), lm(x, lengths=lens), lm(x[:, 1:]), MelSpectrogram().cuda()(x), T.Audio2Mel().cuda()(x.unsqueeze(1))]
st = T.STFT(1024, 256).cuda()
outs += list(st.transform(x)) + [st.magnitude(x)] + list(T.STFTTorchAudio(1024, 300).cuda()(x))
x4 = torch.randn(5, 30000, device="cuda") * 0.1
mel128 = T.LogMelSpectrogram(44100, 128, 2048, 2048, 512).cuda()(x4)
outs.append(mel128)
outs += list(T.STFTTorchAudio(2048, 512).cuda().transform(x4))
big = torch.randn(300, 22050, device="cuda") * 0.1
outs.append(lm(big))
# full-spectrum pair kernel (fmax = Nyquist: no pruning), power / HTK plan, short window (table path), tiny clips
outs.append(T.LogMelSpectrogram(16000, 80, 1024, 1024, 256).cuda()(x))
outs.append(T.LogMelSpectrogramTorchAudio(22050, 64, 1024, 800, 200, -50, 30).cuda()(x))
outs.append(lm(x[:3, :513]))
# waveform-side operators and the DCT
outs.append(PreEmphasis().cuda()(x.unsqueeze(1)))
outs.append(PreEmphasis().cuda()(x[:, 3:4100].unsqueeze(1)))
outs.append(volume_norm_log_torch(x))
outs.append(T.MelToMFCC(40, 80).cuda()(outs[0]))
outs.append(T.MelToMFCC(13, 128).cuda()(mel128))
# round 2: small transform sizes, fused pre-emphasis / frame mask, fused MFCC epilogue, the tcgen05 mel GEMM and the
# tcgen05 STFT kernel (forced on: clips shorter than a batch, edge batches, a batch larger than one wave of CTAs)
import ctypes as C
from pytorch_sound_b200 import _lib
outs += list(T.STFT(512, 128).cuda().transform(x)) + [T.STFT(256, 64, 200).cuda().magnitude(x)]
outs += list(lm(x, lengths=lens, frame_mask=True)) + [lm(x, preemphasis=0.97)]
outs.append(T.MFCC(22050, 80, 1024, 1024, 40, 256, -50, 30, 0., 8000.).cuda()(x))
outs.append(T.MFCC(22050, 80, 1024, 1024, 13, 256, -50, 30, 0., 8000.).cuda()(big))
outs.append(T.LogMelScale(22050, 80, 1024, -50, 30, 0., 8000.).cuda()(st.magnitude(x)))
lib = _lib.lib()
lib.b200mel_debug_set_tc_mode.argtypes = [C.c_int]
lib.b200mel_debug_tc_launch_count.restype = C.c_int64
# SANITIZE_SKIP_TC=1: leave the tcgen05 STFT kernel out.  Its warps hand shared-memory buffers to each other through
# mbarrier arrive (release) / try_wait (acquire); racecheck only models bar.sync / __syncwarp as ordering and reports
# every such hand-off as a hazard, which would bury a real one in the other kernels.
if os.environ.get("SANITIZE_SKIP_TC") != "1":
    lib.b200mel_debug_set_tc_mode(1)
    n0 = lib.b200mel_debug_tc_launch_count()
    outs += [lm(x), lm(x[:, 1:]), lm(x[:3, :513]), lm(big), MelSpectrogram().cuda()(x)]
    assert lib.b200mel_debug_tc_launch_count() == n0 + 5
    lib.b200mel_debug_set_tc_mode(0)
torch.cuda.synchronize()
print("ok", sum(float(o.float().abs().mean()) for o in outs))
