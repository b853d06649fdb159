"""TextAdaptor (ofasys/adaptor/text.py:58-142): token/box/phone/... ids -> embeddings, learned
positions, log-bucketed relative positions; tied output projection."""
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..configure import register_config
from ..module import Embedding
from .base import AdaptorOutput, BaseAdaptor, BaseAdaptorConfig


def make_token_bucket_position(bucket_size, max_position):
    """Relative position -> bucket id (text.py:20-30).  Integer work: must stay bit-exact, including
    the float log/ceil the reference uses (tested against the reference's own table)."""
    context_pos = torch.arange(max_position, dtype=torch.long)[:, None]
    memory_pos = torch.arange(max_position, dtype=torch.long)[None, :]
    relative_pos = context_pos - memory_pos
    sign = torch.sign(relative_pos)
    mid = bucket_size // 2
    abs_pos = torch.where((relative_pos < mid) & (relative_pos > -mid), mid - 1, torch.abs(relative_pos))
    log_pos = torch.ceil(torch.log(abs_pos / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    log_pos = log_pos.int()
    bucket_pos = torch.where(abs_pos.le(mid), relative_pos, log_pos * sign).long()
    return bucket_pos + bucket_size - 1


@dataclass
class TextAdaptorConfig(BaseAdaptorConfig):
    token_bucket_size: int = 256
    share_input_output_embed: bool = True
    output_embed_dim: Optional[int] = 512
    output_dim: Optional[int] = None
    output_bias: bool = False


@register_config("ofasys.adaptor", "text", TextAdaptorConfig)
class TextAdaptor(BaseAdaptor):
    def __init__(self, embed_tokens, dictionary, is_src, general_adaptor, cfg: TextAdaptorConfig):
        super().__init__(embed_tokens, dictionary, is_src, general_adaptor, cfg)
        self.embed_positions = Embedding(cfg.max_position + 2, cfg.embed_dim)
        token_num_rel_dis = 2 * cfg.token_bucket_size - 1
        n_tables = 1 if cfg.share_attn_bias else self.num_layers
        self.token_rel_pos_table_list = nn.ModuleList(
            [Embedding(token_num_rel_dis, cfg.num_attention_heads, zero_init=True) for _ in range(n_tables)]
        )
        self.register_buffer("token_rp_bucket", make_token_bucket_position(cfg.token_bucket_size, cfg.max_position))
        assert cfg.share_input_output_embed, "untied output projection is not on the OFA path"
        self.share_input_output_embed = True
        self.output_dim = cfg.output_dim if cfg.output_dim is not None else len(dictionary)
        self._idx_cache = {}

    def rel_idx(self, T):
        k = (T, self.token_rp_bucket.device)
        if k not in self._idx_cache:
            self._idx_cache[k] = self.token_rp_bucket[:T, :T].to(torch.int32).contiguous()
        return self._idx_cache[k]

    def get_rel_pos_bias(self, batch_size, seq_length, idx, **kwargs):
        """Dense [T, T, H] values as the reference returns them (text.py:101-104); debugging / tests only."""
        return self.token_rel_pos_table_list[idx].weight[self.token_rp_bucket[:seq_length, :seq_length]]

    def forward(self, slot, **kwargs) -> AdaptorOutput:
        tok = slot.value
        pad = self.dictionary.pad()
        masks = tok.eq(pad) if pad is not None else torch.zeros_like(tok, dtype=torch.bool)
        B, T = tok.shape
        # the encoder zeroes padded source rows after the adaptors (model/transformer.py:109-112)
        embed, pos = self.hook(slot, self.embed_positions.weight, tokens=tok, zero_mask=masks if self.is_src else None)
        out = AdaptorOutput(embed, masks, None if pos is None else pos.expand(B, -1, -1), None)
        if self.cfg.use_self_attn_bias:
            out.rel_idx = self.rel_idx(T)
            out.rel_tables = [t.weight for t in self.token_rel_pos_table_list]
        return out

    def forward_output(self, x, extra, slot, **kwargs):
        return self.embed_tokens_T(x), extra
