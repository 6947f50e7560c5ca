"""A few launches of the tensor-core STFT kernel at the C2 shape (for ncu): B200MEL_TC=1 python tools/tc_launch.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_sound_b200.models.transforms import LogMelSpectrogram

m = LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.).cuda()
x = torch.randn(256, 22050, device="cuda") * 0.1
for _ in range(4):
    y = m(x)
torch.cuda.synchronize()
print(float(y.mean()))
