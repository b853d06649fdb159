"""TEST INFRASTRUCTURE ONLY -- seeded parity cases shared by the golden generator, the tests,
smoke() and the bench's CPU leg.  Everything is regenerated from seeds (CPU mt19937 generator), so
fixtures under tests/golden/ only hold *outputs* of the reference, never weights or inputs.

Weights are deliberately NOT the stock init: LayerNorm gains U(.5,1.5), biases
N(0,.1), rel-pos tables N(0,.5), c_attn U(.5,1.5) -- the stock 1/0/0.02 values hide bugs.
"""
from typing import Dict, List, Tuple

import torch

from .oracle_model import AUDIO, BOX, IMAGE, PAD, TEXT, VIDEO, OracleConfig, OSlot

# name -> dict(cfg=..., adaptors=..., inputs spec)
CASES = {
    # small configs whose full logits fit in a fixture
    "text_A": dict(
        cfg=dict(embed_dim=128, heads=2, ffn_dim=512, enc_layers=2, dec_layers=2, vocab=512, mode="A"),
        adaptors=("text",), kind="text", B=3, S=24, T=16,
    ),
    "text_B": dict(
        cfg=dict(embed_dim=128, heads=2, ffn_dim=512, enc_layers=2, dec_layers=2, vocab=512, mode="B"),
        adaptors=("text",), kind="text", B=3, S=24, T=16,
    ),
    "patch_B": dict(
        cfg=dict(embed_dim=128, heads=2, ffn_dim=512, enc_layers=2, dec_layers=2, vocab=512, mode="B"),
        adaptors=("text", "image_patch_embed"), kind="patch", B=2, S=8, T=12,
    ),
    "audio_A": dict(
        cfg=dict(embed_dim=128, heads=2, ffn_dim=512, enc_layers=2, dec_layers=2, vocab=512, mode="A"),
        adaptors=("text", "audio"), kind="audio", B=2, S=6, T=10, L=200,
    ),
    # image_resnet (ResNet-50 as in the reference's `tiny` preset, train-mode BatchNorm) + text, Mode A, 64x64 image
    "resnet_A": dict(
        cfg=dict(embed_dim=128, heads=2, ffn_dim=512, enc_layers=2, dec_layers=2, vocab=512, mode="A", resnet_type="resnet50"),
        adaptors=("text", "image_resnet"), kind="resnet", B=4, S=8, T=12, image=64,
    ),
    # video_image_sequence (16-frame video in configs[4]; here 3 frames of 64x64, one all-zero = padded frame) + text
    "video_A": dict(
        cfg=dict(embed_dim=128, heads=2, ffn_dim=512, enc_layers=2, dec_layers=2, vocab=512, mode="A", resnet_type="resnet50"),
        adaptors=("text", "video_image_sequence"), kind="video", B=2, S=6, T=10, image=64, frames=3,
    ),
    # BASELINE.json configs[0]: text_infilling, OFA-tiny 4L/4L d=256, seq 128, bs 2 (checksums only)
    "cfg1_tiny": dict(
        cfg=dict(embed_dim=256, heads=4, ffn_dim=1024, enc_layers=4, dec_layers=4, vocab=50265, mode="A"),
        adaptors=("text",), kind="text", B=2, S=128, T=128,
    ),
    # BASELINE.json configs[1] at its full model size: image_caption, OFA-base 12L/12L d=768 H=12, 224^2 patch-embed (257
    # tokens) + 8-token prompt -> 64-token caption, V=50265 (bench.py's workload at B=2; checksums only)
    "cfg2_base": dict(
        cfg=dict(embed_dim=768, heads=12, ffn_dim=3072, enc_layers=12, dec_layers=12, vocab=50265, mode="B"),
        adaptors=("text", "image_patch_embed"), kind="patch", B=2, S=8, T=64,
    ),
    # BASELINE.json configs[2] at its full model size: ASR, OFA-base, 10 s fbank [998,80] (ragged lengths) + 12-token
    # prompt -> 128-token transcript, Mode A (audio rel-pos table over 2047 buckets)
    "cfg3_asr_base": dict(
        cfg=dict(embed_dim=768, heads=12, ffn_dim=3072, enc_layers=12, dec_layers=12, vocab=50265, mode="A"),
        adaptors=("text", "audio"), kind="audio", B=2, S=12, T=128, L=998,
    ),
    # OFA-large layer geometry (configs[4]: d=1024, H=16, F=4096) on a shallow 3L/2L stack
    "large_A": dict(
        cfg=dict(embed_dim=1024, heads=16, ffn_dim=4096, enc_layers=3, dec_layers=2, vocab=512, mode="A"),
        adaptors=("text",), kind="text", B=3, S=40, T=24,
    ),
    # ---- BASELINE.json configs[3] / configs[4] at their full model size (oracle/make_golden_full.py; checksums, samples and
    # projections only).  `tasks`: the task batches of ONE step, gradients accumulated over them.
    "cfg4_cotrain_base": dict(
        cfg=dict(embed_dim=768, heads=12, ffn_dim=3072, enc_layers=12, dec_layers=12, vocab=50265, mode="A", resnet_type="resnet101"),
        adaptors=("text", "image_resnet"),
        tasks=[dict(kind="resnet", B=2, S=8, T=64, image=224), dict(kind="resnet", B=2, S=16, T=8, image=224), dict(kind="text", B=2, S=128, T=128)],
    ),
    "cfg5_large_video": dict(
        cfg=dict(embed_dim=1024, heads=16, ffn_dim=4096, enc_layers=24, dec_layers=12, vocab=51265, mode="A", resnet_type="resnet152"),
        adaptors=("text", "image_resnet", "video_image_sequence"),
        tasks=[dict(kind="video", B=1, S=8, T=64, image=224, frames=16)],
    ),
    "cfg5_large_grounding": dict(
        cfg=dict(embed_dim=1024, heads=16, ffn_dim=4096, enc_layers=24, dec_layers=12, vocab=51265, mode="A", resnet_type="resnet152"),
        adaptors=("text", "image_resnet", "video_image_sequence"),
        tasks=[dict(kind="resnet", B=2, S=16, T=5, image=512, box=True)],
    ),
}


def oracle_cfg(name) -> OracleConfig:
    return OracleConfig(**CASES[name]["cfg"])


def synth_tensor(name: str, shape, g: torch.Generator) -> torch.Tensor:
    """Deterministic parameter values by role (role decided from the parameter name)."""
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if "rel_pos_table" in name:
        return torch.randn(shape, generator=g) * 0.5
    if leaf == "c_attn":
        return torch.rand(shape, generator=g) + 0.5
    if ".bn3." in name and leaf == "weight":
        # residual-branch gain U(.1,.3): a random-weight ReLU+BatchNorm ResNet with unit gains is chaotic (perturbations
        # grow ~1.25x per block, so bf16 rounding alone moves the C4 features by >20 %); trained nets are not
        return torch.rand(shape, generator=g) * 0.2 + 0.1
    is_norm = any(t in name for t in ("layer_norm", "layernorm", "_ln.", "attn_ln", ".bn", "downsample.1"))
    if is_norm and leaf == "weight":
        return torch.rand(shape, generator=g) + 0.5
    if is_norm and leaf == "bias":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "running_mean":
        return torch.zeros(shape)
    if leaf == "running_var":
        return torch.ones(shape)
    if leaf == "bias":
        return torch.randn(shape, generator=g) * 0.02
    if "embed_images" in name and leaf == "weight" and len(shape) == 4:  # resnet convs: fan-in scaled
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    if "subsample.conv" in name and leaf == "weight":
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
    if "subsample.out" in name and leaf == "weight":
        return torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
    if "proj.weight" in name and len(shape) == 4:  # patch-embed conv
        return torch.randn(shape, generator=g) * 0.02
    if leaf in ("cls_token", "mask_emb"):
        return torch.randn(shape, generator=g) * 0.02
    # Linear / Embedding weights: BERT init (module/initialize.py:33-36)
    return torch.randn(shape, generator=g) * 0.02


def synth_state_dict(spec: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """spec: name -> shape for every floating-point parameter / BN buffer.  Tied tensors
    (encoder/decoder embed_tokens) are generated once and shared."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(spec):
        if name == "decoder.adaptor.embed_tokens.weight":
            continue
        t = synth_tensor(name, spec[name], g)
        if name == "encoder.adaptor.embed_tokens.weight":
            t[PAD].zero_()  # nn.Embedding(padding_idx) row (module/layer.py:8-15)
        sd[name] = t
    if "encoder.adaptor.embed_tokens.weight" in sd:
        sd["decoder.adaptor.embed_tokens.weight"] = sd["encoder.adaptor.embed_tokens.weight"]
    return sd


def _tokens(g, B, T, V, ragged=True):
    tok = torch.randint(4, V, (B, T), generator=g)
    if ragged and B > 1:  # right-pad the last sequence by ~25% (collate_tokens default)
        cut = T - max(1, T // 4)
        tok[-1, cut:] = PAD
    return tok


def make_inputs(name: str, seed: int = 1234):
    """Returns (slots: List[OSlot], target: LongTensor[B, T])."""
    c = CASES[name]
    g = torch.Generator().manual_seed(seed)
    V = c["cfg"]["vocab"]
    B, S, T = c["B"], c["S"], c["T"]
    slots: List[OSlot] = []
    if c["kind"] == "patch":
        img = torch.randn(B, 3, 224, 224, generator=g)
        slots.append(OSlot(IMAGE, True, img, adaptor="image_patch_embed"))
    if c["kind"] == "resnet":
        img = torch.randn(B, 3, c["image"], c["image"], generator=g)
        slots.append(OSlot(IMAGE, True, img, adaptor="image_resnet"))
    if c["kind"] == "video":
        vid = torch.randn(B, 3, c["frames"], c["image"], c["image"], generator=g)
        vid[-1, :, -1] = 0.0  # the last frame of the last clip is padding (video_image_sequence.py:136-139)
        slots.append(OSlot(VIDEO, True, vid))
    if c["kind"] == "audio":
        L = c["L"]
        fbank = torch.randn(B, L, 80, generator=g)
        lens = torch.tensor([L] + [L - L // 4] * (B - 1), dtype=torch.long)
        slots.append(OSlot(AUDIO, True, {"fbank": fbank, "fbank_lengths": lens}))
    slots.append(OSlot(TEXT, True, _tokens(g, B, S, V)))
    prev = _tokens(g, B, T, V)
    prev[:, 0] = 0  # bos
    slots.append(OSlot(TEXT, False, prev))
    target = torch.roll(prev, -1, dims=1)
    target[:, -1] = 2  # eos
    target[prev == PAD] = PAD
    target[torch.roll(prev == PAD, -1, dims=1)] = PAD
    target[:, -1] = torch.where(prev[:, -1] == PAD, torch.tensor(PAD), torch.tensor(2))
    return slots, target


def make_task_inputs(name: str, seed: int = 1234):
    """[(slots, target)] for the task batches of one step of a multi-task case (`tasks` in CASES)."""
    c = CASES[name]
    out = []
    for ti, t in enumerate(c["tasks"]):
        CASES["_task"] = dict(cfg=c["cfg"], adaptors=c["adaptors"], **t)
        try:
            slots, target = make_inputs("_task", seed + 101 * ti)
        finally:
            del CASES["_task"]
        if t.get("box"):  # BOX target: bos + 4 `<bin>` tokens of the last 1000 vocabulary entries (preprocessor/default/box.py:101-110)
            g = torch.Generator().manual_seed(seed + 101 * ti + 7)
            V, B = c["cfg"]["vocab"], t["B"]
            bins = V - 1000 + torch.randint(0, 1000, (B, 4), generator=g)
            prev = torch.cat([torch.zeros(B, 1, dtype=torch.long), bins], dim=1)
            target = torch.cat([bins, torch.full((B, 1), 2, dtype=torch.long)], dim=1)
            slots = slots[:-1] + [OSlot(BOX, False, prev)]
        out.append((slots, target))
    return out


# ---- full-tensor gradient probes (fixtures hold them instead of 10^8 gradient values) -------------------------------------
GRAD_SAMPLES = 256   # elements at seeded positions per parameter
GRAD_PROJ = 2        # seeded +-1 projections per parameter


def grad_probe_indices(name, numel):
    import zlib

    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randint(0, numel, (1024,), generator=g)[:min(GRAD_SAMPLES, numel)]


def grad_probe_vectors(name, numel):
    import zlib

    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ 0x5BD1E995) & 0x7FFFFFFF)
    return torch.randint(0, 2, (GRAD_PROJ, numel), generator=g, dtype=torch.int8) * 2 - 1


def grad_probes(name, grad):
    """(samples fp32 [<=GRAD_SAMPLES], projections fp64 [GRAD_PROJ]) of one gradient tensor (any device / dtype)."""
    gd = grad.detach().reshape(-1)
    idx = grad_probe_indices(name, gd.numel()).to(gd.device)
    r = grad_probe_vectors(name, gd.numel()).to(gd.device)
    g64 = gd.double()
    return g64[idx].float().cpu(), torch.stack([(g64 * r[i].double()).sum() for i in range(GRAD_PROJ)]).cpu()


def param_spec_from_state_dict(sd) -> Dict[str, Tuple[int, ...]]:
    """Floating-point tensors the synth fills (parameters + BN running stats); integer buckets and
    the `version` buffers are derived, not synthesised."""
    spec = {}
    for k, v in sd.items():
        if not v.is_floating_point() or k.endswith(".version"):
            continue
        spec[k] = tuple(v.shape)
    return spec
