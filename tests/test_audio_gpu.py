"""GPU: Kaldi-compatible fbank + utterance CMVN kernels (csrc/audio.cu) against the oracle
(oracle/oracle_audio.py, pinned to torchaudio's compliance.kaldi.fbank and the reference's UtteranceCMVN by
tests/golden/fbank.pt) and against that fixture.

Tolerance: the kernel's radix-2 fp32 FFT and torch's pocketfft agree to ~1e-6 of the frame's largest bin; a mel energy
far below that level (quiet band next to a loud tone) carries that absolute error, so the log-mel features are compared
as |d| <= 2e-3 + rel-L2 <= 2e-5 per utterance (features are O(10), measured max |d| ~1e-4)."""
import os

import pytest
import torch

from oracle import oracle_audio as oa

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_fbank_batch_ragged_matches_oracle_and_torchaudio_fixture():
    from ofasys_b200.preprocessor.audio import Fbank, utterance_cmvn_

    fx = torch.load(os.path.join(GOLD, "fbank.pt"), weights_only=False)
    wav, lengths = oa.make_case()
    fb = Fbank(num_mel_bins=80, sample_frequency=16000)
    feats, n_frames = fb(wav.cuda(), lengths.cuda())
    torch.cuda.synchronize()
    assert n_frames.tolist() == [f.shape[0] for f in fx["fbank"]]  # integer work: bit-exact
    assert feats.shape == (3, fb.num_frames(wav.shape[1]), 80)
    for b in range(3):
        m = int(n_frames[b])
        ref = oa.fbank(wav[b:b + 1, : int(lengths[b])])
        got = feats[b, :m].cpu()
        assert (got - ref).abs().max().item() <= 2e-3 and _rel(got, ref) <= 2e-5, ((got - ref).abs().max().item(), _rel(got, ref))
        assert (got - fx["fbank"][b]).abs().max().item() <= 2e-3 and _rel(got, fx["fbank"][b]) <= 2e-5
        assert not feats[b, m:].any()  # padding frames are zero
    normed = utterance_cmvn_(feats.clone(), n_frames)
    for b in range(3):
        m = int(n_frames[b])
        ref = torch.from_numpy(oa.utterance_cmvn(oa.fbank(wav[b:b + 1, : int(lengths[b])]).numpy()))
        got = normed[b, :m].cpu()
        # The reference's variance is E[x^2] - mean^2 in float32 (numpy): for a column that hardly varies over the utterance
        # (the bins of the test tone: x ~ 25, var ~ 1e-3) that cancellation leaves no correct digit in ITS result, so the
        # fixture is compared on the columns with var > 0.5 only; the kernel accumulates in double and is held to the
        # float64 evaluation of the same formula on every column.
        feat64 = oa.fbank(wav[b:b + 1, : int(lengths[b])]).double()
        var64 = feat64.var(0, unbiased=False)
        ok = var64 > 0.5
        assert int(ok.sum()) >= 30
        assert (got[:, ok] - ref[:, ok]).abs().max().item() <= 5e-3 and _rel(got[:, ok], ref[:, ok]) <= 2e-4
        assert (got[:, ok] - fx["cmvn"][b][:, ok]).abs().max().item() <= 5e-3
        exact = (feat64 - feat64.mean(0)) / var64.clamp_min(1e-10).sqrt()
        assert _rel(got, exact) <= 1e-4, _rel(got, exact)
        assert abs(got.mean().item()) < 1e-3 and abs(got.std(dim=0, unbiased=False).mean().item() - 1.0) < 1e-3


def test_fbank_full_size_10s_batch_properties():
    """BASELINE configs[2] size: 10 s at 16 kHz -> 998 frames x 80, B = 32.  Size-independent
    properties: shift invariance of the framing (dropping `shift` samples drops exactly the first frame), gain
    (scaling the waveform by a adds 2 ln a to every log-mel energy), and agreement with the oracle on a sample."""
    from ofasys_b200.preprocessor.audio import Fbank

    g = torch.Generator().manual_seed(0)
    B, n = 32, 160000
    wav = (torch.randn(B, n, generator=g) * 1500).cuda()
    fb = Fbank()
    f0, nf = fb(wav)
    assert f0.shape == (B, 998, 80) and nf.tolist() == [998] * B
    f1, _ = fb(wav[:, 160:])
    assert torch.equal(f1[:, :997], f0[:, 1:998])  # same frames, same arithmetic
    f2, _ = fb(wav * 4.0)
    assert (f2 - f0 - 2 * torch.log(torch.tensor(4.0))).abs().max().item() <= 1e-4
    ref = oa.fbank(wav[5:6].cpu())
    assert (f0[5].cpu() - ref).abs().max().item() <= 2e-3 and _rel(f0[5].cpu(), ref) <= 2e-5


def test_fbank_feeds_the_audio_adaptor_contract():
    """Output layout is the AUDIO slot value of the data contract: {'fbank': [B, L, 80] float, 'fbank_lengths': [B] int64}."""
    from ofasys_b200.preprocessor.audio import Fbank

    wav, lengths = oa.make_case()
    feats, n_frames = Fbank()(wav.cuda(), lengths.cuda())
    assert feats.dtype == torch.float32 and n_frames.dtype == torch.int64 and feats.is_contiguous()
    with pytest.raises(Exception):
        Fbank()(wav)  # CPU tensor: no fallback


def test_spec_augment_masks_the_reference_bands():
    """SpecAugment masking on device == the reference transform's own arithmetic (restated in
    oracle_audio.spec_augment, utils/audio_feature_transforms/specaugment.py:79-126) under the same numpy seed, bit-exact for a
    numeric mask value; the utterance-mean fill within fp32 summation order."""
    import numpy as np
    from oracle import oracle_audio as oa
    from ofasys_b200.preprocessor.audio import SpecAugment

    g = torch.Generator().manual_seed(5)
    B, L, F = 3, 120, 80
    feats = torch.randn(B, L, F, generator=g)
    lens = torch.tensor([120, 97, 64])
    for mv in (0.0, None):
        kw = dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=40, time_mask_p=0.2, mask_value=mv)
        np.random.seed(7)
        want = [oa.spec_augment(feats[b, : int(lens[b])].numpy(), **kw) for b in range(B)]
        np.random.seed(7)
        x = feats.clone().cuda()
        SpecAugment(**kw)(x, lens.cuda())
        for b in range(B):
            n = int(lens[b])
            got = x[b, :n].cpu()
            if mv is None:
                assert (got - torch.from_numpy(want[b])).abs().max().item() <= 1e-6
            else:
                assert torch.equal(got, torch.from_numpy(want[b]))
            assert torch.equal(x[b, n:].cpu(), feats[b, n:])  # rows past the utterance untouched
        assert (x.cpu() != feats).any()


def test_image_normalize_and_box_bins_bit_exact():
    """ToTensor + Normalize (preprocessor/default/image.py:110-116) and the `<bin>` quantisation (box.py:101-110) on device,
    bit-exact against the same torch arithmetic the reference runs."""
    from ofasys_b200.preprocessor.image import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, box_to_tokens, normalize_images
    from oracle import oracle_model as om

    g = torch.Generator().manual_seed(3)
    px = torch.randint(0, 256, (2, 37, 53, 3), generator=g, dtype=torch.uint8)
    for mean, std in (((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), (IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD)):
        want = px.permute(0, 3, 1, 2).float().div(255)  # ToTensor
        want = want.sub(torch.tensor(mean).view(1, 3, 1, 1)).div(torch.tensor(std).view(1, 3, 1, 1))  # Normalize
        got = normalize_images(px.cuda(), mean, std)
        assert torch.equal(got.cpu(), want)
    coords = torch.cat([torch.rand(4000, generator=g) * 512, torch.tensor([0.0, 511.0, 512.0, 255.6, 0.2563, 256.2563])])
    want = 50265 + om.quantize_box(coords, 512, 1000)
    assert torch.equal(box_to_tokens(coords.cuda(), 50265).cpu(), want)
