"""TEST INFRASTRUCTURE ONLY (container only: needs /root/reference) -- golden of the reference's own SpecAugmentTransform
(utils/audio_feature_transforms/specaugment.py, imported unmodified through stub packages) on a seeded [97, 80] feature matrix
under np.random.seed(7), for mask_value 0.0 and None (utterance mean).  Output: tests/golden/specaugment.pt.
    python -m oracle.make_golden_specaugment
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import oracle_audio as oa
from . import ref_shim

CASE = dict(time_warp_w=0, freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=40, time_mask_p=0.2)


def case_input():
    g = torch.Generator().manual_seed(5)
    return torch.randn(97, 80, generator=g).numpy()


def main():
    ref_shim.install()
    pk = types.ModuleType("ofasys.utils.audio_feature_transforms")
    pk.__path__ = [os.path.join(ref_shim.REF_PKG, "utils", "audio_feature_transforms")]
    pk.AudioFeatureTransform = type("AudioFeatureTransform", (), {})
    pk.register_audio_feature_transform = lambda name: (lambda c: c)
    sys.modules["ofasys.utils.audio_feature_transforms"] = pk
    mod = importlib.import_module("ofasys.utils.audio_feature_transforms.specaugment")
    x = case_input()
    out = {}
    for tag, mv in (("zero", 0.0), ("mean", None)):
        t = mod.SpecAugmentTransform(CASE["time_warp_w"], CASE["freq_mask_n"], CASE["freq_mask_f"], CASE["time_mask_n"], CASE["time_mask_t"], CASE["time_mask_p"], mv)
        np.random.seed(7)
        ref = t(x)
        np.random.seed(7)
        assert np.array_equal(ref, oa.spec_augment(x, mask_value=mv, **CASE)), tag
        out[tag] = torch.from_numpy(ref.copy())
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "specaugment.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
