"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports /root/reference/pytorch_sound/models/transforms.py,
interface/hifi_gan.py and utils/calculate.py as they are and records their
outputs on seeded inputs.  To make the 2020-era reference importable on this
image (torch 2.11, no librosa) three third-party names are shimmed BEFORE import:

  * `librosa` is absent: a stub module provides `librosa.filters.mel` (the oracle's
    restatement of librosa 0.8.0 — so the filterbank itself is NOT pinned by this
    script, only everything around it) and `librosa.util.pad_center`;
  * `scipy.signal.kaiser` moved to scipy.signal.windows (only imported by the
    reference's PQMF, unused here);
  * `torch.stft` lost its legacy real-valued return: a wrapper restores the
    torch 1.7 behaviour `view_as_real(stft(..., return_complex=True))`;
  * `unidecode` / `inflect` (text cleaners pulled in by settings.py) are stubbed.

Nothing under /root/reference is modified or copied.  The GPU box never runs this.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def install_shims():
    import scipy.signal
    import scipy.signal.windows
    import torch

    from oracle import mel_oracle

    if not hasattr(scipy.signal, "kaiser"):
        scipy.signal.kaiser = scipy.signal.windows.kaiser

    librosa = types.ModuleType("librosa")
    filters = types.ModuleType("librosa.filters")
    util = types.ModuleType("librosa.util")

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney"):
        return mel_oracle.mel_filterbank(sr, n_fft, n_mels, fmin, fmax, htk, norm)

    def pad_center(data, size, axis=-1):
        n = data.shape[axis]
        lpad = int((size - n) // 2)
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (lpad, int(size - n - lpad))
        return np.pad(data, lengths)

    filters.mel = mel
    util.pad_center = pad_center
    librosa.filters = filters
    librosa.util = util
    sys.modules["librosa"] = librosa
    sys.modules["librosa.filters"] = filters
    sys.modules["librosa.util"] = util
    for name in ("unidecode", "inflect"):
        m = types.ModuleType(name)
        m.unidecode = lambda s: s
        m.engine = lambda: None
        sys.modules[name] = m

    real_stft = torch.stft

    def legacy_stft(input, n_fft, hop_length=None, win_length=None, window=None, center=True, pad_mode="reflect",
                    normalized=False, onesided=None, return_complex=None):
        if return_complex is None:
            out = real_stft(input, n_fft, hop_length, win_length, window, center, pad_mode, normalized, onesided,
                            return_complex=True)
            return torch.view_as_real(out)
        return real_stft(input, n_fft, hop_length, win_length, window, center, pad_mode, normalized, onesided,
                         return_complex=return_complex)

    torch.stft = legacy_stft


def main():
    import torch

    install_shims()
    sys.path.insert(0, REF)
    from pytorch_sound.models import transforms as T
    from pytorch_sound.interface.hifi_gan import MelSpectrogram
    from pytorch_sound.utils import calculate
    from pytorch_sound import settings

    from oracle import mel_oracle

    torch.manual_seed(0)
    torch.set_num_threads(1)
    out = {}

    # C1 of BASELINE.json: 1 x 22050 pure sine at the settings.py geometry
    n = np.arange(22050)
    c1 = (0.5 * np.sin(2 * np.pi * 440.0 * n / 22050)).astype(np.float32)[None]
    # seeded sine+noise clips with a length that is not a multiple of hop
    clips = mel_oracle.synth_clips(4, 6000, 22050, seed=20261017)
    noise = np.random.default_rng(7).uniform(-1, 1, size=(3, 4099)).astype(np.float32)

    geo = dict(sample_rate=settings.SAMPLE_RATE, mel_size=settings.MEL_SIZE, n_fft=settings.N_FFT,
               win_length=settings.WIN_LENGTH, hop_length=settings.HOP_LENGTH, mel_min=float(settings.MEL_MIN),
               mel_max=float(settings.MEL_MAX))
    out["settings"] = np.array([settings.SAMPLE_RATE, settings.N_FFT, settings.WIN_LENGTH, settings.HOP_LENGTH,
                                settings.SPEC_SIZE, settings.MEL_SIZE, settings.MEL_MIN, settings.MEL_MAX,
                                settings.MIN_DB, settings.MAX_DB], dtype=np.int64)

    for name, x in (("c1", c1), ("clips", clips), ("noise", noise)):
        xt = torch.from_numpy(x)
        out[f"{name}.wav"] = x
        with torch.no_grad():
            lm = T.LogMelSpectrogram(min_db=settings.MIN_DB, max_db=settings.MAX_DB, **geo)
            out[f"{name}.logmel_clamped"] = lm(xt).numpy()
            lm_nc = T.LogMelSpectrogram(**geo)
            out[f"{name}.logmel"] = lm_nc(xt).numpy()
            out[f"{name}.logmel_off1e-3"] = lm_nc(xt, log_offset=1e-3).numpy()
            st = T.STFT(filter_length=1024, hop_length=256)
            mag, phase = st.transform(xt)
            out[f"{name}.stft_mag"] = mag.numpy()
            out[f"{name}.stft_phase"] = phase.numpy()
            sta = T.STFTTorchAudio(filter_length=1024, hop_length=256)
            re, im = sta(xt)
            out[f"{name}.stfta_re"] = re.numpy()
            out[f"{name}.stfta_im"] = im.numpy()
            mag2, phase2 = sta.transform(xt)
            out[f"{name}.stfta_mag"] = mag2.numpy()
            a2m = T.Audio2Mel()
            out[f"{name}.audio2mel"] = a2m(xt.unsqueeze(1)).numpy()
            hf = MelSpectrogram()
            out[f"{name}.hifi"] = hf(xt).numpy()
            out[f"{name}.norm_mel"] = calculate.norm_mel(lm(xt)).numpy()

    # n_fft = 2048 geometry (BASELINE config C4: 44100 Hz, hop 512, 128 mels)
    x4 = mel_oracle.synth_clips(2, 9000, 44100, seed=20261017 + 4000)
    out["c4.wav"] = x4
    with torch.no_grad():
        lm4 = T.LogMelSpectrogram(sample_rate=44100, mel_size=128, n_fft=2048, win_length=2048, hop_length=512)
        out["c4.logmel"] = lm4(torch.from_numpy(x4)).numpy()
        a2m4 = T.Audio2Mel(n_fft=2048, hop_length=512, win_length=2048, sampling_rate=44100, n_mel_channels=128)
        out["c4.audio2mel"] = a2m4(torch.from_numpy(x4).unsqueeze(1)).numpy()

    # buffers the reference modules register (state_dict compatibility targets)
    out["buf.mel_filter"] = lm.mel_filter.numpy()
    out["buf.hifi_window"] = hf.window.numpy()
    out["buf.stft_window_sq"] = st.square_window.numpy()
    out["kat.db2log"] = np.array([calculate.db2log(np.array(float(settings.MIN_DB))),
                                  calculate.db2log(np.array(float(settings.MAX_DB)))])
    xs = torch.linspace(-1, 1, 11)
    out["kat.unnorm_mel"] = calculate.unnorm_mel(xs).numpy()

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def extras():
    """tests/golden/reference_extra.npz: the torchaudio-wrapping LogMelSpectrogramTorchAudio (run with the
    torchaudio of this image, whose MelSpectrogram defaults equal 0.7.0's: power 2, HTK, no norm), PreEmphasis,
    volume_norm_log(_torch) and MelToMFCC — same shims, same unmodified reference sources."""
    import torch

    sys.path.insert(0, REF)
    from pytorch_sound.models import transforms as T
    from pytorch_sound.models import sound as S
    from pytorch_sound.utils import calculate
    from pytorch_sound import settings

    from oracle import mel_oracle

    torch.set_num_threads(1)
    out = {}
    clips = mel_oracle.synth_clips(4, 6000, 22050, seed=20261017)
    noise = np.random.default_rng(7).uniform(-1, 1, size=(3, 4099)).astype(np.float32)
    with torch.no_grad():
        ta = T.LogMelSpectrogramTorchAudio(settings.SAMPLE_RATE, settings.MEL_SIZE, settings.N_FFT, settings.WIN_LENGTH,
                                           settings.HOP_LENGTH, settings.MIN_DB, settings.MAX_DB, float(settings.MEL_MIN),
                                           float(settings.MEL_MAX))
        ta_w = T.LogMelSpectrogramTorchAudio(22050, 64, 1024, 800, 200, -50, 30)  # win < n_fft, f_max default
        out["buf.ta_fb"] = ta.melfunc.mel_scale.fb.numpy()
        out["buf.ta_window"] = ta.melfunc.spectrogram.window.numpy()
        pe = S.PreEmphasis()
        out["buf.flipped_filter"] = pe.flipped_filter.numpy()
        mf = T.MelToMFCC(settings.MFCC_SIZE, settings.MEL_SIZE)
        out["buf.dct_mat"] = mf.dct_mat.numpy()
        lm = T.LogMelSpectrogram(settings.SAMPLE_RATE, settings.MEL_SIZE, settings.N_FFT, settings.WIN_LENGTH,
                                 settings.HOP_LENGTH, settings.MIN_DB, settings.MAX_DB, float(settings.MEL_MIN),
                                 float(settings.MEL_MAX))
        for name, x in (("clips", clips), ("noise", noise)):
            xt = torch.from_numpy(x)
            out[f"{name}.wav"] = x
            out[f"{name}.ta_logmel"] = ta(xt).numpy()
            out[f"{name}.ta_logmel_win800"] = ta_w(xt).numpy()
            out[f"{name}.preemphasis"] = pe(xt.unsqueeze(1)).numpy()
            out[f"{name}.volume_norm_torch"] = calculate.volume_norm_log_torch(xt).numpy()
            out[f"{name}.volume_norm_np"] = calculate.volume_norm_log(x)
            out[f"{name}.mfcc"] = mf(lm(xt)).numpy()
        # other geometries of BASELINE.json, run by the reference itself: C5 (16 kHz, fmax = Nyquist) through
        # LogMelSpectrogram / Audio2Mel / hifi MelSpectrogram, and a hop that is not a divisor of n_fft
        from pytorch_sound.interface.hifi_gan import MelSpectrogram
        x5 = mel_oracle.synth_clips(3, 8000, 16000, seed=20261017 + 5000)
        out["c5.wav"] = x5
        x5t = torch.from_numpy(x5)
        out["c5.logmel_clamped"] = T.LogMelSpectrogram(16000, 80, 1024, 1024, 256, -50, 30, 0.0, 8000.0)(x5t).numpy()
        out["c5.audio2mel"] = T.Audio2Mel(sampling_rate=16000)(x5t.unsqueeze(1)).numpy()
        out["c5.hifi"] = MelSpectrogram(sampling_rate=16000, fmax=7600.)(x5t).numpy()
        out["clips.logmel_hop300"] = T.LogMelSpectrogram(22050, 80, 1024, 1024, 300, None, None, 0.0, 8000.0)(
            torch.from_numpy(clips)).numpy()
        st = T.STFTTorchAudio(filter_length=1024, hop_length=200, win_length=800, n_fft=1024)
        re, im = st(torch.from_numpy(clips))
        out["clips.stfta_win800_re"], out["clips.stfta_win800_im"] = re.numpy(), im.numpy()
    path = os.path.join(HERE, "reference_extra.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def round2():
    """tests/golden/reference_round2.npz: the reference's multi_stft_loss (models/sound.py:106-147), its STFT /
    LogMelSpectrogram at transform sizes below 1024, and SpectrogramMasker (models/transforms.py:397-416).

    Two more shims, both for code that hard-codes `.cuda()` in the reference (models/sound.py:113-117,
    models/transforms.py:405-406) and therefore cannot run on this GPU-less container as written:
    `torch.nn.Module.cuda` is a no-op while this function runs.  The arithmetic is untouched."""
    import torch

    sys.path.insert(0, REF)
    real_cuda = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, device=None: self
    try:
        from pytorch_sound.models import transforms as T
        from pytorch_sound.models import sound as S

        from oracle import mel_oracle

        torch.set_num_threads(1)
        out = {}
        target = mel_oracle.synth_clips(3, 16000, 22050, seed=20261017 + 7000)
        rng = np.random.default_rng(11)
        pred = (0.9 * target + 0.02 * rng.standard_normal(target.shape)).astype(np.float32)
        out["loss.pred"], out["loss.target"] = pred, target
        params = [(1024, 600, 120), (2048, 1200, 240), (512, 240, 50)]
        out["loss.params"] = np.array(params, dtype=np.int64)
        with torch.no_grad():
            tot, sc, mag = S.multi_stft_loss(torch.from_numpy(pred), torch.from_numpy(target), params)
            out["loss.values"] = np.array([float(tot), float(sc), float(mag)], dtype=np.float64)
            tot2, sc2, mag2 = S.multi_stft_loss(torch.from_numpy(pred), torch.from_numpy(target), [(512, 512, 128)], eps=1e-3)
            out["loss.values_512"] = np.array([float(tot2), float(sc2), float(mag2)], dtype=np.float64)
            for fft, win, hop in params:
                st = S.STFT(win, hop, win, fft)  # = STFTTorchAudio(filter_length=win, hop_length=hop, win_length=win, n_fft=fft)
                out[f"loss.target_mag_{fft}"] = st.transform(torch.from_numpy(target[:1]))[0].numpy()

            clips = mel_oracle.synth_clips(4, 6000, 22050, seed=20261017)
            out["clips.wav"] = clips
            xt = torch.from_numpy(clips)
            for n, hop in ((512, 128), (256, 64), (128, 100)):
                mag_, ph_ = T.STFT(filter_length=n, hop_length=hop).transform(xt)
                out[f"clips.stft{n}_mag"], out[f"clips.stft{n}_phase"] = mag_.numpy(), ph_.numpy()
            re, im = T.STFTTorchAudio(filter_length=400, hop_length=160, win_length=400, n_fft=512)(xt)
            out["clips.stfta_win400_fft512_re"], out["clips.stfta_win400_fft512_im"] = re.numpy(), im.numpy()
            x16 = mel_oracle.synth_clips(3, 8000, 16000, seed=20261017 + 5000)
            out["c16.wav"] = x16
            out["c16.logmel512"] = T.LogMelSpectrogram(16000, 40, 512, 512, 128, -50, 30, 0.0, 8000.0)(torch.from_numpy(x16)).numpy()
            out["c16.audio2mel512"] = T.Audio2Mel(n_fft=512, hop_length=128, win_length=512, sampling_rate=16000,
                                                  n_mel_channels=40)(torch.from_numpy(x16).unsqueeze(1)).numpy()

            # SpectrogramMasker on pad_collate_fn-style masks (ones over the valid samples)
            for win, hop, L, lens in ((1024, 256, 22050, [22050, 1, 700, 12345, 21800]), (800, 200, 6000, [6000, 4500, 5999, 399])):
                mask = np.zeros((len(lens), L), dtype=np.float32)
                for i, n in enumerate(lens):
                    mask[i, :n] = 1
                out[f"masker.{win}_{hop}.lengths"] = np.array(lens, dtype=np.int64)
                out[f"masker.{win}_{hop}.L"] = np.array(L)
                out[f"masker.{win}_{hop}.out"] = T.SpectrogramMasker(win, hop)(torch.from_numpy(mask)).numpy()
    finally:
        torch.nn.Module.cuda = real_cuda
    path = os.path.join(HERE, "reference_round2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if "--round2-only" in sys.argv:
        install_shims()
        round2()
    else:
        if "--extras-only" not in sys.argv:
            main()
        else:
            install_shims()
        extras()
        round2()
