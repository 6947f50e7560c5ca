"""A few launches of the spectrum-output kernel at the C2 shape (for ncu): python tools/spec_launch.py [mag|transform|reim]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_sound_b200.models import transforms as T

kind = sys.argv[1] if len(sys.argv) > 1 else "mag"
x = torch.randn(256, 22050, device="cuda") * 0.1
st = T.STFT(filter_length=1024, hop_length=256).cuda()
sa = T.STFTTorchAudio(filter_length=1024, hop_length=256, win_length=1024).cuda()
for _ in range(4):
    y = st.magnitude(x) if kind == "mag" else (st.transform(x)[0] if kind == "transform" else sa(x)[0])
torch.cuda.synchronize()
print(float(y.mean()))
