"""MultiheadAttention with the reference's constructor, parameter names and numerics
(ofasys/module/multihead_attention.py:21-353), executed by csrc/gemm.cu + csrc/attn.cu.

Reproduced quirks:
  * attn_bias is None and not static_kv  -> the reference short-circuits to
    F.multi_head_attention_forward: scale head_dim**-0.5 and c_attn is ignored (:155-186);
  * otherwise scale (head_dim * scale_factor)**-0.5, additive bias, causal mask, -inf on padded keys,
    fp32 softmax, per-head c_attn before out_proj (:188-353).
The dense [B*H, T, S] attn_bias of the reference is replaced by ops.PositionBias (the same numbers,
generated inside the attention kernel instead of being materialised per layer).
"""
import math

import torch
import torch.nn as nn

from .. import ops


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, kdim=None, vdim=None, dropout=0.0, bias=True, add_bias_kv=False,
                 add_zero_attn=False, self_attention=False, encoder_decoder_attention=False, scale_factor=2,
                 scale_heads=False, use_fused=True):
        super().__init__()
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self.qkv_same_dim = self.kdim == embed_dim and self.vdim == embed_dim
        self.num_heads = num_heads
        self.dropout_p = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        assert self.head_dim == 64, "ofasys_b200 attention kernels are built for head_dim 64 (every OFA preset)"
        assert not add_bias_kv and not add_zero_attn, "add_bias_kv / add_zero_attn are not on the OFA path"
        self.scale_factor = scale_factor
        self.scaling = float(self.head_dim * scale_factor) ** -0.5
        self.self_attention = self_attention
        self.encoder_decoder_attention = encoder_decoder_attention
        self.c_attn = nn.Parameter(torch.ones((num_heads,)), requires_grad=True) if scale_heads else None
        assert not self.self_attention or self.qkv_same_dim
        self.k_proj = nn.Linear(self.kdim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(self.vdim, embed_dim, bias=bias)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        gain = 1 / math.sqrt(2) if self.qkv_same_dim else 1.0
        for p in (self.k_proj, self.v_proj, self.q_proj):
            nn.init.xavier_uniform_(p.weight, gain=gain)
        nn.init.xavier_uniform_(self.out_proj.weight)
        if self.out_proj.bias is not None:
            nn.init.constant_(self.out_proj.bias, 0.0)

    def _cat(self, names, attr):
        ts = [getattr(getattr(self, n), attr) for n in names]
        return None if ts[0] is None else torch.cat(ts, dim=0)

    def forward(self, query, key=None, value=None, key_padding_mask=None, incremental_state=None, need_weights=False,
                static_kv=False, attn_mask=None, need_head_weights=False, attn_bias=None, batch_first=False, causal=None):
        """query/key: T x B x C as in the reference, or B x T x C with batch_first=True (internal layout).
        attn_bias: None | False | ops.PositionBias.  attn_mask: only the causal future mask is supported;
        pass causal=True (a non-None attn_mask is taken to be that mask).  Returns (out bf16, None)."""
        if incremental_state is not None or need_weights or need_head_weights:
            raise NotImplementedError("incremental decoding / returned attention maps are outside the fwd+bwd hot path")
        if isinstance(attn_bias, torch.Tensor):
            raise NotImplementedError("dense attn_bias tensors are replaced by ofasys_b200.ops.PositionBias")
        if causal is None:
            causal = attn_mask is not None
        if not batch_first:
            query = query.transpose(0, 1)
            key = None if key is None or key is query else key.transpose(0, 1)
        x = ops.to_bf16(query)
        fast = attn_bias is None and not static_kv  # reference :150-186
        scale = float(self.head_dim) ** -0.5 if fast else self.scaling
        bias = attn_bias if isinstance(attn_bias, ops.PositionBias) else None
        H = self.num_heads
        drop = ops.dropout_state(x.device).spec(self.dropout_p) if self.training else None  # on the probabilities (:335)
        if self.self_attention or key is None or key is query:
            qkv = ops.linear(x, self._cat(("q_proj", "k_proj", "v_proj"), "weight"), self._cat(("q_proj", "k_proj", "v_proj"), "bias"))
            ctx = ops.attention(qkv, None, H, scale, bias, key_padding_mask, causal, drop=drop)
        else:
            mem = ops.to_bf16(key)
            q = ops.linear(x, self.q_proj.weight, self.q_proj.bias)
            kv = ops.linear(mem, self._cat(("k_proj", "v_proj"), "weight"), self._cat(("k_proj", "v_proj"), "bias"))
            ctx = ops.attention(q, kv, H, scale, bias, key_padding_mask, causal, drop=drop)
        w_out = self.out_proj.weight
        if self.c_attn is not None and not fast:
            w_out = ops.scale_cols(w_out, self.c_attn, self.head_dim)  # einsum('tbhd,h->tbhd') folded into out_proj (:342-346)
        out = ops.linear(ctx, w_out, self.out_proj.bias)
        if not batch_first:
            out = out.transpose(0, 1)
        return out, None
