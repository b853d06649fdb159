"""OFAGeneralAdaptor (ofasys/adaptor/general.py:49-316): dispatch slots to modality adaptors in
ModalityType order, concatenate along time, build the position terms of the attention bias."""
import dataclasses
from typing import Any, Dict, List

import torch
import torch.nn as nn

from .. import ops
from ..configure import BaseDataclass, ConfigStore
from ..module import Embedding
from ..preprocessor import ModalityType, Slot
from .base import AdaptorOutput, BaseAdaptor
from . import text, image_resnet, image_patch_embed, audio, video_image_sequence  # noqa: F401  (registers the adaptors)

_ORDER = ["text", "image_resnet", "image_patch_embed", "audio_fbank", "video_image_sequence"]


def _make_adaptor_config():
    store = ConfigStore()
    names = [n for n in _ORDER if store.contain("ofasys.adaptor", n)]
    names += sorted(n for n in store.names("ofasys.adaptor") if n not in names)
    flds = [(n, type(store.get("ofasys.adaptor", n).config), dataclasses.field(default_factory=type(store.get("ofasys.adaptor", n).config)))
            for n in names]
    return dataclasses.make_dataclass("OFAAdaptorConfig", flds, bases=(BaseDataclass,))


OFAAdaptorConfig = _make_adaptor_config()

default_adaptor = {
    ModalityType.TEXT: "text",
    ModalityType.IMAGE: "image_resnet",
    ModalityType.BOX: "text",
    ModalityType.AUDIO: "audio_fbank",
    ModalityType.PHONE: "text",
    ModalityType.VIDEO: "video_image_sequence",
    ModalityType.MOTION: "text",
    ModalityType.STRUCT: "text",
    ModalityType.CATEGORY: "text",
}


class OFAGeneralAdaptor(nn.Module):
    _embed_tokens = None  # shared by the encoder's and the decoder's adaptor (general.py:191-221)

    def __init__(self, cfg, dictionary, is_src):
        super().__init__()
        self.embed_tokens = self.build_embedding(cfg, dictionary)
        self.cfg = cfg
        self.is_src = is_src
        self.name2adaptor: Dict[str, BaseAdaptor] = {}
        for f in dataclasses.fields(cfg.adaptor):
            if f.name.startswith("_"):
                continue
            if f.name in ("image_resnet", "video_image_sequence", "image_vit") and not is_src:
                continue
            config = getattr(cfg.adaptor, f.name)
            config.parse_from_model_cfg(cfg)
            if not config.is_active:
                continue
            target = ConfigStore().get("ofasys.adaptor", f.name).target
            self.name2adaptor[f.name] = target(self.embed_tokens, dictionary, is_src, self, config)
            setattr(self, f.name, self.name2adaptor[f.name])
        embed_dim = cfg.encoder.embed_dim if is_src else cfg.decoder.embed_dim
        self.num_attention_heads = cfg.encoder.attention_heads if is_src else cfg.decoder.attention_heads
        self.pos_scaling = float(embed_dim / cfg.encoder.attention_heads * cfg.attn_scale_factor) ** -0.5
        if not cfg.entangle_position_embedding:
            self.pos_q_linear = nn.Linear(embed_dim, embed_dim)
            self.pos_k_linear = nn.Linear(embed_dim, embed_dim)
        self._idx_cache = {}
        self.dense_bias = True  # False: structured position terms inside the mma.sync attention kernels (incremental decoding)

    def build_embedding(self, cfg, dictionary):
        if OFAGeneralAdaptor._embed_tokens is not None:
            return OFAGeneralAdaptor._embed_tokens
        assert cfg.share_all_embeddings
        assert cfg.encoder.embed_dim == cfg.decoder.embed_dim
        assert cfg.max_source_positions == cfg.max_target_positions
        emb = Embedding(len(dictionary), cfg.encoder.embed_dim, padding_idx=dictionary.pad())
        cfg.share_decoder_input_output_embed = True
        if getattr(cfg, "freeze_encoder_embedding", False):
            emb.weight.requires_grad = False
        OFAGeneralAdaptor._embed_tokens = emb
        return emb

    def get_adaptor(self, slot: Slot) -> BaseAdaptor:
        name = slot.get_attr("adaptor")
        return self.name2adaptor[name if name else default_adaptor[slot.modality]]

    def forward(self, slots: List[Slot], **kwargs):
        outs = [None] * len(slots)
        cnt = 0
        for mod in ModalityType:  # fixed order, as the reference (general.py:137-149)
            for i, slot in enumerate(slots):
                if slot.modality == mod:
                    outs[i] = self.get_adaptor(slot)(slot, **kwargs)
                    cnt += 1
            if cnt == len(slots):
                break
        assert cnt == len(slots), cnt
        out = self.concat(outs)
        return out.embed, out.masks, out.pos_embed, out.self_attn_bias, None

    def forward_output(self, x, extra: Dict[str, Any], slots: List[Slot], **kwargs):
        output_slot = None
        for slot in slots:
            if not slot.is_src:
                assert output_slot is None, "supports only one target slot"
                output_slot = slot
        assert output_slot
        return self.get_adaptor(output_slot).forward_output(x, extra, slot=output_slot)

    COMPACT_ABOVE = 1024  # concatenated table rows above which the bucket ids are compacted to the used ones

    def _global_idx(self, outs: List[AdaptorOutput]):
        """int32 [S, S]: block-diagonal bucket ids, each slot's ids offset into the concatenated table;
        -1 where the reference adds no relative bias (general.py:270-280).  Cached per shape signature."""
        key = tuple((o.seq_length, None if o.rel_idx is None else (o.rel_idx.data_ptr(), o.rel_tables[0].shape[0])) for o in outs)
        if key in self._idx_cache:
            return self._idx_cache[key]
        S = sum(o.seq_length for o in outs)
        dev = outs[0].embed.device
        idx = torch.full((S, S), -1, dtype=torch.int32, device=dev)
        start, off = 0, 0
        for o in outs:
            T = o.seq_length
            if o.rel_idx is not None:
                idx[start:start + T, start:start + T] = o.rel_idx + off
                off += o.rel_tables[0].shape[0]
            start += T
        used = None
        if off > self.COMPACT_ABOVE:
            # the attention kernels stage one table column per head in shared memory: keep only the rows this shape
            # can address (e.g. 729 of the image table's 6892 for a 14x14 grid)
            used = torch.unique(idx[idx >= 0].long())
            lut = torch.full((off,), -1, dtype=torch.int32, device=dev)
            lut[used] = torch.arange(used.numel(), dtype=torch.int32, device=dev)
            idx = torch.where(idx >= 0, lut[idx.clamp_min(0).long()], idx)
        self._idx_cache[key] = (idx, used)
        return idx, used

    def concat(self, outs: List[AdaptorOutput]) -> AdaptorOutput:
        one = len(outs) == 1
        embed = outs[0].embed if one else torch.cat([o.embed for o in outs], dim=1)
        masks = outs[0].masks if one else torch.cat([o.masks for o in outs], dim=1)
        B = embed.shape[0]
        if any(o.pos_embed is None for o in outs):
            pos = None
        else:
            pos = outs[0].pos_embed if one else torch.cat([o.pos_embed[:1] for o in outs], dim=1).expand(B, -1, -1)
        out = AdaptorOutput(embed, masks, pos, None)
        if not self.cfg.use_self_attn_bias:
            return out
        # abs term (general.py:223-243): batch-invariant projections, scale applied inside the kernel
        p1 = pos[:1]
        pq = ops.linear(p1, self.pos_q_linear.weight, self.pos_q_linear.bias)
        pk = ops.linear(p1, self.pos_k_linear.weight, self.pos_k_linear.bias)
        num_layers = self.cfg.encoder.layers if self.is_src else self.cfg.decoder.layers
        n_tables = 1 if self.cfg.share_attn_bias else num_layers
        with_rel = [o for o in outs if o.rel_idx is not None]
        idx, used = self._global_idx(outs) if with_rel else (None, None)
        # dense abs-pos term [H, S, S], once per forward and batch-invariant (the reference expands it to [B, H, S, S] and
        # clones it per layer, general.py:223-282); every layer's attention adds its own table gather to it in ONE tile
        abs_t = ops.abs_pos(pq, pk, self.num_attention_heads) if self.dense_bias else None
        biases = []
        for l in range(n_tables):
            table = None
            if with_rel:
                tabs = [o.rel_tables[l] for o in with_rel]
                table = tabs[0] if len(tabs) == 1 else torch.cat(tabs, dim=0)
                if used is not None:
                    table = table.index_select(0, used)
            biases.append(ops.PositionBias(pq, pk, idx, table, abs=abs_t))
        out.self_attn_bias = biases
        return out

    def update_sample(self, sample):
        for a in self.name2adaptor.values():
            a.update_sample(sample)
        return sample
