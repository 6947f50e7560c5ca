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
torch.cuda.synchronize()
print("ok", sum(float(o.float().abs().mean()) for o in outs))
