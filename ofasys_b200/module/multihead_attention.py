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

    # ---- incremental decoding (multihead_attention.py:188-279: prev_key / prev_value / prev_key_padding_mask) ----------
    # State per attention module inside the caller's `incremental_state` dict: a K|V cache [B, capacity, 2d] (bf16,
    # K in columns [0, d), V in [d, 2d) -- the layout the attention kernel reads with strides, so nothing is re-packed
    # per step), its length, and the key-padding mask of the cached keys.  Self-attention appends the new token's K|V;
    # encoder-decoder attention (static_kv) projects the encoder output once.
    @property
    def _state_key(self):
        return f"ofab_attn_state.{id(self)}"

    def _forward_incremental(self, query, key, key_padding_mask, incremental_state, static_kv, attn_bias, batch_first):
        if not batch_first:
            query = query.transpose(0, 1)
            key = None if key is None or key is query else key.transpose(0, 1)
        x = ops.to_bf16(query)
        B, Tq, d = x.shape
        H = self.num_heads
        st = incremental_state.setdefault(self._state_key, {})
        bias = attn_bias if isinstance(attn_bias, ops.PositionBias) else None
        # the reference takes the F.multi_head_attention_forward shortcut only when incremental_state is None
        # (multihead_attention.py:155-160): with a cache it is always the manual path -- (head_dim * scale_factor)**-0.5, c_attn
        fast = False
        scale = self.scaling
        if static_kv:
            assert self.encoder_decoder_attention and not self.self_attention
            if "kv" not in st:  # first step: project the encoder output once (:188-196)
                mem = ops.to_bf16(key)
                st["kv"] = ops.linear(mem, self._cat(("k_proj", "v_proj"), "weight"), self._cat(("k_proj", "v_proj"), "bias")).detach()
                st["len"] = st["kv"].shape[1]
                st["kpm"] = None if key_padding_mask is None else key_padding_mask.to(torch.uint8).contiguous()
            q = ops.linear(x, self.q_proj.weight, self.q_proj.bias)
            kv, kpm = st["kv"], st["kpm"]
            causal = False
        else:
            assert self.self_attention
            qkv = ops.linear(x, self._cat(("q_proj", "k_proj", "v_proj"), "weight"), self._cat(("q_proj", "k_proj", "v_proj"), "bias"))
            q = qkv[..., :d]
            L0 = st.get("len", 0)
            assert Tq == 1 or L0 == 0, "cached steps feed one new position at a time (the first call may carry a prefix)"
            L = L0 + Tq
            cache = st.get("kv")
            if cache is None or cache.shape[1] < L or cache.shape[0] != B:
                cap = max(16, 2 * L)
                new = torch.empty((B, cap, 2 * d), dtype=torch.bfloat16, device=x.device)
                mask = torch.zeros((B, cap), dtype=torch.uint8, device=x.device)
                if cache is not None and L0 > 0:
                    new[:, :L0].copy_(cache[:, :L0])
                    mask[:, :L0].copy_(st["kpm_buf"][:, :L0])
                st["kv"], st["kpm_buf"], cache = new, mask, new
            cache[:, L0:L].copy_(qkv[..., d:].detach())  # torch.cat([prev_key, k]) (:241-256) as an in-place append
            if key_padding_mask is not None:
                st["kpm_buf"][:, L0:L].copy_(key_padding_mask.to(torch.uint8))
                st["has_kpm"] = True
            st["len"] = L
            kv = cache[:, :L]
            kpm = st["kpm_buf"][:, :L].contiguous() if st.get("has_kpm") else None
            causal = Tq > 1  # a prefix fed at once; a single new position sees every cached key
        if bias is not None:  # the caller passes the bias rows of the new positions (transformer.py:474-475)
            assert bias.pq is None or bias.pq.shape[1] == Tq
        ctx = ops.attention(q, kv, H, scale, bias, kpm, causal)
        w_out = self.out_proj.weight
        if self.c_attn is not None and not fast:
            w_out = ops.scale_cols(w_out, self.c_attn, self.head_dim)
        out = ops.linear(ctx, w_out, self.out_proj.bias)
        if not batch_first:
            out = out.transpose(0, 1)
        return out, None

    def reorder_incremental_state(self, incremental_state, new_order):
        """Beam reordering of this module's cache (multihead_attention.py:393-409)."""
        st = incremental_state.get(self._state_key)
        if st is None:
            return incremental_state
        for k in ("kv", "kpm_buf", "kpm"):
            if st.get(k) is not None:
                st[k] = st[k].index_select(0, new_order)
        return incremental_state

    def _cat(self, names, attr):
        """The projections' weights (or biases) as ONE packed operand: the parameters share a storage (ops.pack_params), so
        nothing is concatenated per forward and the packed gradient is handed back as slices (the reference computes three
        separate F.linear calls, multihead_attention.py:199-218)."""
        ts = [getattr(getattr(self, n), attr) for n in names]
        if ts[0] is None:
            return None
        if not ts[0].is_cuda:
            return torch.cat(ts, dim=0)
        return ops.pack_params(ts)

    def forward(self, query, key=None, value=None, key_padding_mask=None, incremental_state=None, need_weights=False,
                static_kv=False, attn_mask=None, need_head_weights=False, attn_bias=None, batch_first=False, causal=None):
        """query/key: T x B x C as in the reference, or B x T x C with batch_first=True (internal layout).
        attn_bias: None | False | ops.PositionBias.  attn_mask: only the causal future mask is supported;
        pass causal=True (a non-None attn_mask is taken to be that mask).  Returns (out bf16, None)."""
        if need_weights or need_head_weights:
            raise NotImplementedError("attention maps are never materialised by the fused kernel")
        if incremental_state is not None:
            return self._forward_incremental(query, key, key_padding_mask, incremental_state, static_kv, attn_bias, batch_first)
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
