"""GPU audio front end (SURVEY 8f "next" #4): Kaldi-compatible log-mel filterbank + utterance CMVN.

Replaces, at the entry of the model path, the reference's CPU feature extraction in the data loader
(`ofasys/preprocessor/default/audio.py:283-305` -> `_get_torchaudio_fbank` `:507-516` ->
`torchaudio.compliance.kaldi.fbank(waveform, num_mel_bins=80, sample_frequency=16000)`; then
`ofasys/utils/audio_feature_transforms/utterance_cmvn.py:33-44`).  The result feeds `AudioFbankAdaptor` as
`{'fbank': [B, L, 80], 'fbank_lengths': [B]}` (SURVEY 8b data contract).

Host side (here): the window and the mel matrix, computed once with the published formulas (torchaudio 2.11
`compliance/kaldi.py`: `_feature_window_function` povey = hann(periodic=False) ** 0.85; `get_mel_banks` without VTLN).
Device side: `csrc/audio.cu` (one CTA per frame).  No CPU fallback.
"""
import ctypes
import math

import torch

from .. import _lib


def _next_pow2(n):
    return 1 if n == 0 else 2 ** (n - 1).bit_length()


def kaldi_window(win: int) -> torch.Tensor:
    """povey window, fp32 [win]."""
    return torch.hann_window(win, periodic=False, dtype=torch.float32).pow(0.85)


def kaldi_mel_banks(n_mel: int, nfft: int, sample_rate: float, low_freq: float = 20.0, high_freq: float = 0.0) -> torch.Tensor:
    """fp32 [n_mel, nfft/2 + 1]: triangular filters equally spaced on the mel scale 1127 ln(1 + f/700) between low_freq and
    high_freq (<= 0: offset from Nyquist), evaluated at the FFT bin centres k * sample_rate / nfft; the Nyquist column is 0."""
    nyquist = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyquist
    mel = lambda f: 1127.0 * math.log(1.0 + f / 700.0)
    mel_low, mel_high = mel(low_freq), mel(high_freq)
    delta = (mel_high - mel_low) / (n_mel + 1)
    b = torch.arange(n_mel).unsqueeze(1)
    left, center, right = mel_low + b * delta, mel_low + (b + 1.0) * delta, mel_low + (b + 2.0) * delta
    bin_mel = (1127.0 * (1.0 + (sample_rate / nfft) * torch.arange(nfft / 2) / 700.0).log()).unsqueeze(0)
    up, down = (bin_mel - left) / (center - left), (right - bin_mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down)).to(torch.float32)
    return torch.nn.functional.pad(bins, (0, 1))


class Fbank:
    """fbank = Fbank(num_mel_bins=80, sample_frequency=16000)(waveform[B, n] fp32 cuda, lengths[B] or None)
    -> (features fp32 [B, max_frames, n_mel], n_frames int64 [B]).  Defaults = torchaudio.compliance.kaldi.fbank's."""

    def __init__(self, num_mel_bins=80, sample_frequency=16000.0, frame_length=25.0, frame_shift=10.0, preemphasis_coefficient=0.97,
                 low_freq=20.0, high_freq=0.0):
        self.n_mel = int(num_mel_bins)
        self.win = int(sample_frequency * frame_length * 0.001)
        self.shift = int(sample_frequency * frame_shift * 0.001)
        self.nfft = _next_pow2(self.win)
        self.preemph = float(preemphasis_coefficient)
        self._window = kaldi_window(self.win)
        self._melT = kaldi_mel_banks(self.n_mel, self.nfft, float(sample_frequency), low_freq, high_freq).t().contiguous()
        self._dev = {}

    def num_frames(self, n_samples: int) -> int:
        return 0 if n_samples < self.win else 1 + (n_samples - self.win) // self.shift

    def __call__(self, waveform: torch.Tensor, lengths: torch.Tensor = None):
        if not waveform.is_cuda:
            raise _lib.OfabError("ofasys_b200 Fbank needs a CUDA waveform (no CPU fallback)")
        wav = waveform.to(torch.float32)
        wav = wav if wav.stride(-1) == 1 else wav.contiguous()
        B, n = wav.shape
        dev = wav.device
        if dev not in self._dev:
            self._dev[dev] = (self._window.to(dev), self._melT.to(dev))
        window, melT = self._dev[dev]
        max_frames = max(1, self.num_frames(n))
        out = torch.empty((B, max_frames, self.n_mel), dtype=torch.float32, device=dev)
        n_frames = torch.empty(B, dtype=torch.int64, device=dev)
        if lengths is not None:
            lengths = lengths.to(device=dev, dtype=torch.int64).contiguous()
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        _lib.call("ofab_fbank", p(wav), wav.stride(0), p(lengths), B, n, p(window), p(melT), self.win, self.shift, self.nfft, self.n_mel,
                  self.preemph, p(out), p(n_frames), max_frames, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out, n_frames


def utterance_cmvn_(feats: torch.Tensor, n_frames: torch.Tensor = None, norm_means=True, norm_vars=True):
    """In place on feats fp32 [B, max_frames, n_feat] over the first n_frames[b] rows of each utterance."""
    if not feats.is_cuda or feats.dtype != torch.float32 or not feats.is_contiguous():
        raise _lib.OfabError("utterance_cmvn_ needs a contiguous CUDA fp32 tensor (no CPU fallback)")
    B, T, F = feats.shape
    nf = None if n_frames is None else n_frames.to(device=feats.device, dtype=torch.int64).contiguous()
    _lib.call("ofab_utterance_cmvn", ctypes.c_void_p(feats.data_ptr()), None if nf is None else ctypes.c_void_p(nf.data_ptr()), B, T, F,
              int(norm_means), int(norm_vars), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return feats


class SpecAugment:
    """The training-time masking of `SpecAugmentTransform` (utils/audio_feature_transforms/specaugment.py:79-126) on device
    features [B, max_frames, n_feat]: `freq_mask_n` bands of width < `freq_mask_f`, `time_mask_n` bands of width <
    min(time_mask_t, floor(n_frames * time_mask_p)), filled with `mask_value` (None: the utterance mean).  The random draws are
    made on the host with numpy's global generator in the reference's own order, one utterance after the other, so a run seeded
    like the reference masks the same bands.  `time_warp_w` (cv2 resize) is not implemented."""

    def __init__(self, time_warp_w=0, freq_mask_n=0, freq_mask_f=0, time_mask_n=0, time_mask_t=0, time_mask_p=0.0, mask_value=0.0):
        if time_warp_w:
            raise NotImplementedError("SpecAugment time warping (cv2 resize) is not implemented on the device path")
        self.freq_mask_n, self.freq_mask_f = int(freq_mask_n), int(freq_mask_f)
        self.time_mask_n, self.time_mask_t, self.time_mask_p = int(time_mask_n), int(time_mask_t), float(time_mask_p)
        self.mask_value = mask_value

    def draw(self, n_frames, n_freqs):
        """host: the (start, width) pairs one call of the reference transform would draw for one utterance"""
        import numpy as np

        bands = []
        if n_frames == 0 or n_freqs < self.freq_mask_f:
            return [(0, 0)] * (self.freq_mask_n + self.time_mask_n)
        for _ in range(self.freq_mask_n):
            f = np.random.randint(0, self.freq_mask_f)
            f0 = np.random.randint(0, n_freqs - f)
            bands.append((f0, f))
        max_t = min(self.time_mask_t, math.floor(n_frames * self.time_mask_p))
        for _ in range(self.time_mask_n):
            if max_t < 1:
                bands.append((0, 0))
                continue
            t = np.random.randint(0, max_t)
            t0 = np.random.randint(0, n_frames - t)
            bands.append((t0, t))
        return bands

    def __call__(self, feats, n_frames=None):
        if not feats.is_cuda or feats.dtype != torch.float32 or not feats.is_contiguous():
            raise _lib.OfabError("SpecAugment needs a contiguous fp32 CUDA tensor [B, max_frames, n_feat] (no CPU fallback)")
        B, L, F = feats.shape
        lens = [L] * B if n_frames is None else [int(v) for v in n_frames.tolist()]
        table = torch.tensor([self.draw(min(n, L), F) for n in lens], dtype=torch.int32).reshape(B, -1, 2).to(feats.device)
        nf = None if n_frames is None else n_frames.to(device=feats.device, dtype=torch.int64).contiguous()
        use_mean = self.mask_value is None
        scratch = torch.empty(B, dtype=torch.float32, device=feats.device) if use_mean else None
        _lib.call("ofab_spec_augment", ctypes.c_void_p(feats.data_ptr()), None if nf is None else ctypes.c_void_p(nf.data_ptr()), B, L, F,
                  ctypes.c_void_p(table.data_ptr()), self.freq_mask_n, self.time_mask_n, 0.0 if use_mean else float(self.mask_value), int(use_mean),
                  None if scratch is None else ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        return feats
