"""Parity at the FULL per-GPU shard shapes of BASELINE.json's multi-GPU configs:

    C3  256 x 88 200   (batch 2048 x 4 s over 8 GPUs)          n_fft 1024, hop 256, 80 mels, fmax 8000
    C4   16 x 441 000  (batch 128 x 10 s @44.1 kHz over 8 GPUs) n_fft 2048, hop 512, 128 mels, fmax sr/2
    C5 8192 x 8 000    (one streamed batch of the 1M-clip run)  n_fft 1024, hop 256, 80 mels, sr 16000

Each shape: float64 oracle on the first 64 clips (C4: all 16), oracle on EVERY clip's edge frames, and the
size-independent properties (clip independence / determinism bit for bit, batch permutation equivariance).
Needs a B200: `-m gpu`."""
import numpy as np
import pytest

from oracle import mel_oracle as mo

pytestmark = pytest.mark.gpu

TOL = 1e-4  # |y - y_ref64| <= 1e-4 * max(1, |y_ref64|), north-star / SURVEY 8d

SHAPES = {
    "C3": dict(B=256, L=88200, n_oracle=64, geo=dict(sample_rate=22050, mel_size=80, n_fft=1024, win_length=1024,
                                                    hop_length=256, mel_min=0.0, mel_max=8000.0), cfg=3),
    "C4": dict(B=16, L=441000, n_oracle=16, geo=dict(sample_rate=44100, mel_size=128, n_fft=2048, win_length=2048,
                                                    hop_length=512, mel_min=0.0, mel_max=None), cfg=4),
    "C5": dict(B=8192, L=8000, n_oracle=64, geo=dict(sample_rate=16000, mel_size=80, n_fft=1024, win_length=1024,
                                                    hop_length=256, mel_min=0.0, mel_max=8000.0), cfg=5),
}


def synth(B, L, sr, seed):
    """SURVEY 8d waveforms; batches beyond 256 clips tile the 72 sinusoids with fresh noise per clip."""
    if B <= 256:
        return mo.synth_clips(B, L, sr, seed=seed)
    out = np.empty((B, L), dtype=np.float32)
    for i in range(0, B, 256):
        n = min(256, B - i)
        out[i:i + n] = mo.synth_clips(n, L, sr, seed=seed + i, first_clip=i)
    return out


def edge_frames_oracle(x, geo, n_edge=3):
    """float64 log-mel of the first and last `n_edge` frames of every clip, from short head / tail segments.

    A centred frame t covers samples [t*hop - n/2, t*hop + n/2): the first frames only see the clip's head, the
    last ones only its tail, so the oracle runs on (n_edge*hop + 2n)-sample segments cut on the frame grid
    instead of on the whole (B, L) array."""
    n, hop = geo["n_fft"], geo["hop_length"]
    L = x.shape[1]
    T = 1 + L // hop
    seg = n_edge * hop + 2 * n
    head = mo.log_mel_spectrogram(x[:, :seg], **geo, clamp=False)[:, :, :n_edge]
    q = max(0, (L - seg) // hop)           # tail segment starts on the frame grid, at sample q*hop
    tail_full = mo.log_mel_spectrogram(x[:, q * hop:], **geo, clamp=False)
    assert tail_full.shape[2] == T - q
    return head, tail_full[:, :, -n_edge:]


@pytest.mark.parametrize("name", ["C3", "C4", "C5"])
def test_full_shard_shape(name):
    import torch

    from pytorch_sound_b200.models import transforms as T

    assert torch.cuda.is_available()
    s = SHAPES[name]
    B, L, geo = s["B"], s["L"], s["geo"]
    hop = geo["hop_length"]
    x = synth(B, L, geo["sample_rate"], seed=20261017 + 1000 * s["cfg"])
    xg = torch.from_numpy(x).cuda()
    lm = T.LogMelSpectrogram(**geo).cuda()
    y = lm(xg)
    Tn = 1 + L // hop
    assert y.shape == (B, geo["mel_size"], Tn) and y.dtype == torch.float32 and bool(torch.isfinite(y).all())

    # (1) clip independence + determinism, bit for bit
    for i in sorted({0, 1, B // 2, B - 1}):
        assert torch.equal(lm(xg[i:i + 1])[0], y[i]), f"clip {i} alone differs from clip {i} in the batch"
    assert torch.equal(lm(xg), y)
    # (2) batch permutation equivariance
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    assert torch.equal(lm(xg[perm]), y[perm])

    # (3) float64 oracle on the first clips, every frame
    n = s["n_oracle"]
    yh = y[:n].cpu().numpy()
    chunk = 8 if L > 100000 else 64
    err = 0.0
    for i in range(0, n, chunk):
        ref = mo.log_mel_spectrogram(x[i:i + chunk], **geo, clamp=False)
        err = max(err, mo.parity_error(yh[i:i + chunk], ref))
    print(f"{name}: first {n} clips max err {err:.2e}")
    assert err < TOL

    # (4) float64 oracle on the edge frames of EVERY clip
    ya = y.cpu().numpy()
    e_head = e_tail = 0.0
    for i in range(0, B, 1024):
        head, tail = edge_frames_oracle(x[i:i + 1024], geo)
        e_head = max(e_head, mo.parity_error(ya[i:i + 1024, :, :3], head))
        e_tail = max(e_tail, mo.parity_error(ya[i:i + 1024, :, -3:], tail))
    print(f"{name}: edge frames of all {B} clips: head {e_head:.2e}, tail {e_tail:.2e}")
    assert e_head < TOL and e_tail < TOL

    # (5) the settings.py clamp on the same batch only clamps (checksum of the clamped tensor vs clamped y)
    lo, hi = float(mo.db2log(-50)), float(mo.db2log(30))
    yc = T.LogMelSpectrogram(min_db=-50, max_db=30, **geo).cuda()(xg)
    assert torch.equal(yc, y.clamp(lo, hi))
