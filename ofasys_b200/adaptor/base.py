"""Adaptor plugin API (ofasys/adaptor/base.py:19-266): AdaptorOutput, BaseAdaptorConfig, BaseAdaptor.

Differences from the reference that are visible through the API (DESIGN.md "boundary"):
  * `self_attn_bias` holds structured position terms instead of dense [B, H, T, T] tensors: each
    adaptor reports its relative-position bucket ids (`rel_idx`, int32 [T, T]) and per-layer tables
    (`rel_tables`); OFAGeneralAdaptor.concat turns them into one ops.PositionBias per layer.  The
    numbers are those of get_rel_pos_bias / expand_rel_pos_bias (base.py:183-189,242-258).
  * `pos_embed` is batch-invariant and returned as an expanded view [B, T, d] of [1, T, d].
The forward hook (base.py:152-191: embed_scale, +pos if entangled, +type if src, LayerNorm(s)) is
fused into one kernel per adaptor (csrc/embed_ce.cu) and therefore called explicitly via `hook()`.
"""
import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..configure import BaseDataclass
from ..module import Embedding, LayerNorm


@dataclass
class AdaptorOutput:
    embed: torch.Tensor  # B x T x d  (fp32 residual stream)
    masks: torch.Tensor  # B x T bool, True = padding
    pos_embed: Optional[torch.Tensor]  # B x T x d (expanded view of 1 x T x d)
    self_attn_bias: Any  # List[ops.PositionBias] after concat; None before
    modal_mask: Optional[torch.Tensor] = None
    rel_idx: Optional[torch.Tensor] = None  # int32 [T, T] local bucket ids of this slot (None: no relative bias)
    rel_tables: Optional[List[torch.Tensor]] = None  # per layer [n_buckets, H]

    def __post_init__(self):
        assert self.embed is not None
        b, t, h = self.embed.shape
        if self.masks is not None:
            assert self.masks.shape == (b, t)
        if self.pos_embed is not None:
            assert self.pos_embed.shape == (b, t, h)

    @property
    def seq_length(self):
        return self.embed.shape[1]


@dataclass
class BaseAdaptorConfig(BaseDataclass):
    is_active: bool = False
    layernorm_embedding: bool = True
    layernorm_position: bool = True
    add_type_embedding: bool = True
    entangle_position_embedding: bool = False
    no_scale_embedding: bool = True
    scale_embedding_gradient: float = 1.0
    dropout: Optional[float] = None
    embed_dim: Optional[int] = None
    num_attention_heads: Optional[int] = None
    encoder_layers: Optional[int] = None
    decoder_layers: Optional[int] = None
    max_position: Optional[int] = None
    use_self_attn_bias: Optional[bool] = None
    share_attn_bias: Optional[bool] = None

    def parse_from_model_cfg(self, model_cfg):
        """base.py:83-101: inherit unset fields from the model config."""
        def pick(cur, new):
            return new if cur is None else cur

        self.dropout = pick(self.dropout, model_cfg.dropout)
        self.embed_dim = pick(self.embed_dim, model_cfg.encoder.embed_dim)
        self.num_attention_heads = pick(self.num_attention_heads, model_cfg.encoder.attention_heads)
        self.encoder_layers = pick(self.encoder_layers, model_cfg.encoder.layers)
        self.decoder_layers = pick(self.decoder_layers, model_cfg.decoder.layers)
        self.max_position = pick(self.max_position, model_cfg.max_source_positions)
        self.use_self_attn_bias = pick(self.use_self_attn_bias, model_cfg.use_self_attn_bias)
        self.share_attn_bias = pick(self.share_attn_bias, model_cfg.share_attn_bias)
        self.entangle_position_embedding = pick(self.entangle_position_embedding, model_cfg.entangle_position_embedding)


class BaseAdaptor(nn.Module):
    def __init__(self, embed_tokens, dictionary, is_src: bool, general_adaptor, cfg: BaseAdaptorConfig):
        super().__init__()
        self._embed_tokens = [embed_tokens]  # not registered as a child (shared, owned by the general adaptor)
        self.dictionary = dictionary
        self.is_src = is_src
        self._general_adaptor = [general_adaptor]
        self.cfg = cfg
        self.num_layers = cfg.encoder_layers if is_src else cfg.decoder_layers
        if cfg.dropout:
            self.dropout_p = cfg.dropout
        else:
            self.dropout_p = 0.0
        assert cfg.layernorm_embedding and cfg.layernorm_position, "ofasys_b200 implements the default layernorm_embedding/position=True"
        assert cfg.no_scale_embedding and cfg.scale_embedding_gradient == 1.0, "embed_scale != 1 is not on the OFA path"
        self.layernorm_embedding = LayerNorm(cfg.embed_dim)
        self.layernorm_position = LayerNorm(cfg.embed_dim)
        self.type_embedding = Embedding(1, cfg.embed_dim) if cfg.add_type_embedding else None
        self.embed_scale = 1.0 if cfg.no_scale_embedding else math.sqrt(cfg.embed_dim)

    @property
    def general_adaptor(self):
        return self._general_adaptor[0]

    @property
    def embed_weight(self):
        return self._embed_tokens[0].weight

    def embed_tokens(self, x):
        raise NotImplementedError("token gathers are fused into hook(); use hook(tokens=...)")

    def embed_tokens_T(self, x):
        """tied output projection F.linear(x, embed_tokens.weight) (base.py:131)."""
        return ops.linear(ops.to_bf16(x), self.embed_weight, None)

    def hook(self, slot, pos_table, tokens=None, dense=None, cls=None, zero_mask=None):
        """Fused forward_hook_fn (base.py:152-191) -> (embed fp32 [B,T,d], pos_embed bf16 [1,T,d] or None).
        pos_table: bf16 rows [>=T, d] indexed by position t."""
        drop = ops.dropout_state(self.layernorm_embedding.weight.device).spec(self.dropout_p) if self.training else None
        entangle = bool(self.cfg.entangle_position_embedding)
        type_vec = self.type_embedding.weight if (slot.is_src and self.type_embedding is not None) else None
        pad = self.dictionary.pad() if tokens is not None and self.dictionary is not None else None
        embed = ops.embed_ln(
            self.layernorm_embedding.weight, self.layernorm_embedding.bias, tokens=tokens,
            E=self.embed_weight if tokens is not None else None, dense=dense, cls=cls,
            pos=pos_table if entangle else None, type_vec=type_vec, zero_mask=zero_mask,
            eps=self.layernorm_embedding.eps, padding_idx=pad, drop=drop,
        )
        T = embed.shape[1]
        pos = None
        if not entangle and pos_table is not None:
            pos = self.layernorm_position(pos_table[:T]).unsqueeze(0)  # LN(pos) feeds only the attention biases
        return embed, pos

    def forward(self, slot, **kwargs) -> AdaptorOutput:
        raise NotImplementedError

    def forward_output(self, x, extra: Dict[str, Any], slot, **kwargs):
        return x, extra

    def get_rel_pos_bias(self, batch_size, seq_length, idx, **kwargs):
        raise NotImplementedError

    def upgrade_state_dict_named(self, state_dict, name):
        pass

    def update_sample(self, sample):
        return sample
