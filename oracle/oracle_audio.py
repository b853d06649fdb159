"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the audio front end the reference runs in its data loader
(SURVEY 8f next #4).  Only tests/ may import it.

The algorithm lives in a third-party dependency that is not under /root/reference: `torchaudio.compliance.kaldi.fbank`
(torchaudio 2.11.0 is what this image pins; requirements.txt of the reference pins nothing).  Call site:
ofasys/preprocessor/default/audio.py:507-516 `ta_kaldi.fbank(waveform, num_mel_bins=n_bins, sample_frequency=sample_rate)`
-- every other argument at its default.  This file restates that published algorithm (function names of
torchaudio/compliance/kaldi.py in the comments) and ofasys/utils/audio_feature_transforms/utterance_cmvn.py:33-44.
Pinning: tests/test_oracle_golden.py checks it against torchaudio itself (installed here and on the GPU box) and
against tests/golden/fbank.pt (torchaudio's output + the reference's own UtteranceCMVN, oracle/make_golden_audio.py).
"""
import math

import numpy as np
import torch

EPS = torch.finfo(torch.float32).eps  # _get_epsilon


def _mel(freq):  # mel_scale / mel_scale_scalar
    return 1127.0 * math.log(1.0 + freq / 700.0)


def mel_banks(num_bins, padded, sample_freq, low_freq=20.0, high_freq=0.0):
    """get_mel_banks with vtln_warp = 1.0 -> [num_bins, padded / 2]."""
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    lo, hi = _mel(low_freq), _mel(high_freq)
    delta = (hi - lo) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left, center, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + fft_bin_width * torch.arange(padded / 2) / 700.0).log()).unsqueeze(0)
    up, down = (mel - left) / (center - left), (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down))


def fbank(waveform, num_mel_bins=80, sample_frequency=16000.0, frame_length=25.0, frame_shift=10.0, preemph=0.97):
    """waveform [1, n] (channel 0 is used) -> [m, num_mel_bins] fp32; kaldi.py fbank with its defaults."""
    wav = waveform[0].to(torch.float32)
    shift = int(sample_frequency * frame_shift * 0.001)  # _get_waveform_and_window_properties
    win = int(sample_frequency * frame_length * 0.001)
    padded = 1 if win == 0 else 2 ** (win - 1).bit_length()
    n = wav.numel()
    if n < win:
        return torch.empty((0, num_mel_bins))
    m = 1 + (n - win) // shift  # _get_strided, snip_edges
    frames = wav.as_strided((m, win), (shift, 1))
    frames = frames - frames.mean(dim=1, keepdim=True)  # remove_dc_offset
    prev = torch.nn.functional.pad(frames.unsqueeze(0), (1, 0), mode="replicate").squeeze(0)[:, :-1]
    frames = frames - preemph * prev  # preemphasis
    frames = frames * torch.hann_window(win, periodic=False).pow(0.85).unsqueeze(0)  # povey
    frames = torch.nn.functional.pad(frames, (0, padded - win))
    spec = torch.fft.rfft(frames).abs().pow(2.0)  # use_power
    mel = torch.nn.functional.pad(mel_banks(num_mel_bins, padded, sample_frequency).to(torch.float32), (0, 1))
    return torch.max(torch.mm(spec, mel.T), torch.tensor(EPS)).log()  # use_log_fbank


def utterance_cmvn(x: np.ndarray, norm_means=True, norm_vars=True) -> np.ndarray:
    """utterance_cmvn.py:33-44 on [frames, features]."""
    mean = x.mean(axis=0)
    square_sums = (x ** 2).sum(axis=0)
    if norm_means:
        x = np.subtract(x, mean)
    if norm_vars:
        var = square_sums / x.shape[0] - mean ** 2
        std = np.sqrt(np.maximum(var, 1e-10))
        x = np.divide(x, std)
    return x


def make_case(seed=11, B=3, n=16000 * 2 + 123):
    """Seeded int16-scaled waveforms (as load_waveform(normalization=False) yields them) with ragged lengths."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 16000.0
    wav = torch.stack([3000.0 * torch.sin(2 * math.pi * (220.0 * (i + 1)) * t) + 800.0 * torch.randn(n, generator=g) for i in range(B)])
    wav[1, 20000:] *= 0.01  # a quiet tail
    lengths = torch.tensor([n, n - 7000, 5000][:B], dtype=torch.long)
    return wav, lengths


def spec_augment(spectrogram, time_warp_w=0, freq_mask_n=0, freq_mask_f=0, time_mask_n=0, time_mask_t=0, time_mask_p=0.0, mask_value=0.0):
    """utils/audio_feature_transforms/specaugment.py:79-126 (SpecAugmentTransform.__call__, no time warping): numpy in, numpy
    out; draws from numpy's global generator in the reference's order."""
    import math

    import numpy as np

    assert time_warp_w == 0
    distorted = spectrogram.copy()
    num_frames, num_freqs = spectrogram.shape
    if mask_value is None:
        mask_value = spectrogram.mean()
    if num_frames == 0 or num_freqs < freq_mask_f:
        return spectrogram
    for _ in range(freq_mask_n):
        f = np.random.randint(0, freq_mask_f)
        f0 = np.random.randint(0, num_freqs - f)
        if f != 0:
            distorted[:, f0:f0 + f] = mask_value
    max_t = min(time_mask_t, math.floor(num_frames * time_mask_p))
    if max_t < 1:
        return distorted
    for _ in range(time_mask_n):
        t = np.random.randint(0, max_t)
        t0 = np.random.randint(0, num_frames - t)
        if t != 0:
            distorted[t0:t0 + t, :] = mask_value
    return distorted
