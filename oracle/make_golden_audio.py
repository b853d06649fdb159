"""TEST INFRASTRUCTURE ONLY (container only) -- golden vectors for the audio front end: torchaudio's own
compliance.kaldi.fbank (the third-party function the reference calls, audio.py:507-516) and the reference's own
UtteranceCMVN (utils/audio_feature_transforms/utterance_cmvn.py, imported unmodified) on oracle_audio.make_case().
Output: tests/golden/fbank.pt.      python -m oracle.make_golden_audio
"""
import importlib
import os
import sys
import types

import torch

from . import oracle_audio as oa
from . import ref_shim


def _reference_cmvn():
    ref_shim.install()
    pkg = types.ModuleType("ofasys.utils.audio_feature_transforms")
    pkg.__path__ = [os.path.join(ref_shim.REF_PKG, "utils", "audio_feature_transforms")]
    pkg.AudioFeatureTransform = object
    pkg.register_audio_feature_transform = lambda name: (lambda cls: cls)
    sys.modules["ofasys.utils.audio_feature_transforms"] = pkg
    return importlib.import_module("ofasys.utils.audio_feature_transforms.utterance_cmvn").UtteranceCMVN


def main():
    import torchaudio.compliance.kaldi as ta_kaldi

    cmvn = _reference_cmvn()()
    wav, lengths = oa.make_case()
    feats, normed = [], []
    for b in range(wav.shape[0]):
        f = ta_kaldi.fbank(wav[b:b + 1, : int(lengths[b])], num_mel_bins=80, sample_frequency=16000)  # audio.py:513
        feats.append(f)
        normed.append(torch.from_numpy(cmvn(f.numpy())))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fbank.pt")
    torch.save({"fbank": [f.to(torch.float32) for f in feats], "cmvn": [x.to(torch.float32) for x in normed]}, path)
    print("wrote", path, os.path.getsize(path), "bytes; frames", [f.shape[0] for f in feats])


if __name__ == "__main__":
    main()
