"""Stage-by-stage check of the tensor-core STFT kernel (stft_tc.cuh) on a GPU: the debug taps of batch 0 against the
numpy model (tests/tc_model.py), the magnitude tap against numpy.fft, the log-mel output against the CUDA-core kernel
and the float64 oracle; then timings of both kernels.  `python tools/tc_debug.py [B] [L]`."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tc_model as tm  # noqa: E402
from oracle import mel_oracle as mo  # noqa: E402
from pytorch_sound_b200 import _lib  # noqa: E402
from pytorch_sound_b200.models import transforms as T  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 22050
    lib = _lib.lib()
    lib.b200mel_debug_set_tc_taps.argtypes = [C.c_void_p] * 4
    lib.b200mel_debug_set_tc_taps.restype = None
    lib.b200mel_debug_set_tc_mode.argtypes = [C.c_int]
    lib.b200mel_debug_tc_launch_count.restype = C.c_int64
    wav = mo.synth_clips(B, L, 22050, seed=3).astype(np.float32)
    x = torch.from_numpy(wav).cuda()
    mod = T.LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0.0, 8000.0).cuda()
    lib.b200mel_debug_set_tc_mode(0)
    y_ref = mod(x)
    torch.cuda.synchronize()
    Tn = y_ref.shape[-1]
    mag = torch.zeros(B, 384, Tn, device="cuda")
    d1 = torch.zeros(256, 32, device="cuda")
    d2 = torch.zeros(128, 48, device="cuda")
    lib.b200mel_debug_set_tc_taps(mag.data_ptr(), d1.data_ptr(), d2.data_ptr(), None)
    lib.b200mel_debug_set_tc_mode(1)
    n0 = lib.b200mel_debug_tc_launch_count()
    y = mod(x)
    torch.cuda.synchronize()
    print("tc launches:", lib.b200mel_debug_tc_launch_count() - n0, "frames", Tn)
    lib.b200mel_debug_set_tc_taps(None, None, None, None)

    # model of batch 0 = clip 0, frames 0..7
    padded = mo.reflect_pad(wav[0].astype(np.float64), 512).astype(np.float32)
    taps = {}
    tm.group_magnitudes(padded[:tm.SPAN], emulate=True, taps=taps)
    D1 = taps["D1"].reshape(256, 32)          # [(n2, t), col]
    g1 = d1.cpu().numpy().astype(np.float64)
    s1 = np.abs(D1).max()
    print("stage-1 accumulators: max |err| / max = %.3e   (max %.3e)" % (np.abs(g1 - D1).max() / s1, s1))
    if np.abs(g1 - D1).max() / s1 > 1e-5:
        bad = np.argwhere(np.abs(g1 - D1) > 1e-5 * s1)
        print("  first bad entries (row = n2 * 8 + t, col):", bad[:10].tolist())
        print("  got", g1[tuple(bad[0])], "want", D1[tuple(bad[0])])
        print("  rows with errors:", np.unique(bad[:, 0])[:40].tolist())
        print("  cols with errors:", np.unique(bad[:, 1])[:40].tolist())
    D2 = taps["D2"].reshape(128, 48)
    g2 = d2.cpu().numpy().astype(np.float64)
    s2 = np.abs(D2).max()
    print("stage-2 accumulators: max |err| / max = %.3e   (max %.3e)" % (np.abs(g2 - D2).max() / s2, s2))
    if np.abs(g2 - D2).max() / s2 > 1e-5:
        bad = np.argwhere(np.abs(g2 - D2) > 1e-5 * s2)
        print("  first bad entries (row = slot * 8 + t, col):", bad[:10].tolist())
        print("  rows with errors:", np.unique(bad[:, 0])[:40].tolist())
        print("  cols with errors:", np.unique(bad[:, 1])[:48].tolist())
    ref = np.abs(mo.stft_complex(wav.astype(np.float64), 1024, 256))[:, :384, :]
    gm = mag.cpu().numpy().astype(np.float64)
    print("magnitudes vs float64: max |err| / max = %.3e" % (np.abs(gm - ref).max() / ref.max()))
    per_clip = np.abs(gm - ref).max(axis=(1, 2)) / ref.max()
    print("  per clip:", ["%.1e" % v for v in per_clip[:8]])
    per_frame = np.abs(gm[0] - ref[0]).max(axis=0) / ref.max()
    print("  clip 0 per frame:", ["%.0e" % v for v in per_frame])
    y64 = mo.log_mel_spectrogram(wav.astype(np.float64), min_db=-50, max_db=30, mel_max=8000.0)
    print("log-mel parity vs float64 oracle: tensor-core %.3e   cuda-core %.3e   (tolerance 1e-4)"
          % (mo.parity_error(y.cpu().numpy(), y64), mo.parity_error(y_ref.cpu().numpy(), y64)))
    print("tensor-core vs cuda-core kernel: max abs diff %.3e" % (y - y_ref).abs().max().item())

    # timing at the C2 shape
    xb = torch.randn(256, 22050, device="cuda") * 0.1
    for mode in (0, 1):
        lib.b200mel_debug_set_tc_mode(mode)
        for _ in range(5):
            mod(xb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            mod(xb)
        e1.record()
        torch.cuda.synchronize()
        print("C2 (256 x 22050) mode %d: %.2f us per call (eager loop, includes launch overhead)" % (mode, e0.elapsed_time(e1) * 20))


if __name__ == "__main__":
    main()
