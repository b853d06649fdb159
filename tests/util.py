"""Shared helpers for the parity tests."""
import os

import torch

import ofasys_b200 as ob
from oracle import cases
from oracle import oracle_model as om

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TTS_ONLY = ("prenet", "postnet", "feat_proj", "eos_proj", "mask_emb")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def load_golden(name):
    return torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)


def build_product(name, device=None, dtype=None):
    """The ofasys_b200 model configured like oracle case `name` (weights NOT loaded)."""
    c = cases.CASES[name]
    cc = c["cfg"]
    cfg = ob.GeneralistModelConfig.default()
    cfg.dropout = 0.0
    cfg.attention_dropout = 0.0
    if cc["mode"] == "B":
        cfg.use_self_attn_bias = False
        cfg.entangle_position_embedding = True
    m = ob.GeneralistModel(cfg)
    d, h, f = cc["embed_dim"], cc["heads"], cc["ffn_dim"]
    m.cfg.encoder.embed_dim = m.cfg.decoder.embed_dim = d
    m.cfg.encoder.ffn_embed_dim = m.cfg.decoder.ffn_embed_dim = f
    m.cfg.decoder.input_dim = m.cfg.decoder.output_dim = d
    m.cfg.encoder.attention_heads = m.cfg.decoder.attention_heads = h
    m.cfg.encoder.layers, m.cfg.decoder.layers = cc["enc_layers"], cc["dec_layers"]
    if "resnet_type" in cc:
        m.cfg.adaptor.image_resnet.resnet_type = cc["resnet_type"]
    for a in c["adaptors"]:
        a = "audio_fbank" if a == "audio" else a
        acfg = getattr(m.cfg.adaptor, a)
        acfg.is_active = True
        if cc["mode"] == "B":
            acfg.entangle_position_embedding = True
        if a == "image_patch_embed":
            acfg.embed_dim = d
        if a == "image_resnet" and "resnet_type" in cc:
            acfg.resnet_type = cc["resnet_type"]
    if cc["mode"] == "B":
        m.cfg.adaptor.text.entangle_position_embedding = True
    m.initialize(ob.Dictionary(n_dummy=cc["vocab"] - 4))
    if dtype is not None:
        m = m.to(dtype)
    if device is not None:
        m = m.to(device)
    return m


def to_product_slots(slots, device, float_dtype=None):
    out = []
    for s in slots:
        v = s.value
        if isinstance(v, dict):
            v = {k: (t.to(device) if not t.is_floating_point() or float_dtype is None else t.to(device, float_dtype)) for k, t in v.items()}
        else:
            v = v.to(device) if (not v.is_floating_point() or float_dtype is None) else v.to(device, float_dtype)
        out.append(ob.Slot(ob.ModalityType(s.modality), s.is_src, v, attributes=f"adaptor={s.adaptor}" if s.adaptor else None))
    return out


def bf16_round_state_dict(sd):
    """The weights the bf16 product actually holds, as fp32 (so the oracle sees identical parameters)."""
    out = {}
    seen = {}
    for k, v in sd.items():
        if v.is_floating_point():
            key = v.data_ptr()
            if key not in seen:
                seen[key] = v.to(torch.bfloat16).float()
            out[k] = seen[key]
        else:
            out[k] = v
    return out
