"""Encoder / decoder layer blocks with the reference's structure and parameter names
(ofasys/module/transformer_layer.py:18-209, 212-495): pre-LN, NormFormer extras (attn_ln,
ffn_layernorm, c_attn) on by default.  The residual stream is fp32 [B, T, C]; GEMM operands are bf16.

Fused execution of one encoder layer (8 kernels + 1 attention):
  LN -> QKV GEMM -> attention -> out_proj GEMM -> [attn_ln, +residual, final_layer_norm] ->
  fc1 GEMM -> [GELU, ffn_layernorm] -> fc2 GEMM (+bias +residual epilogue)
"""
import torch
import torch.nn as nn

from .. import ops
from .layer_norm import LayerNorm
from .multihead_attention import MultiheadAttention
from .transformer_config import TransformerConfig


def _check(cfg, side):
    if not getattr(cfg, side).normalize_before:
        raise NotImplementedError("ofasys_b200 implements the reference default pre-LN layers (normalize_before=True)")
    if cfg.activation_fn != "gelu":
        raise NotImplementedError("only activation_fn='gelu' (config/default_model.yaml) is on the hot path")
    if getattr(cfg, "modal_ffn", False) or getattr(cfg, "scale_resids", False):
        raise NotImplementedError("modal_ffn / scale_resids are off in every BASELINE config and not implemented")


def _branch_drop(layer, x):
    """Descriptor of the block's residual-branch masks: `dropout_module` then `drop_path` on the branch before it is
    added to the residual (transformer_layer.py:181,87,203 / :433,466,489,333).  None in eval mode or when both are 0.
    x: [B, T, C] (drop-path drops whole samples = groups of T rows)."""
    if not layer.training:
        return None
    return ops.dropout_state(x.device).spec(layer.dropout_p, layer.drop_path_rate, x.shape[1])


def _ffn(layer, x2):
    """x2 = final_layer_norm(x) bf16 -> fc2(ffn_ln(act_drop(gelu(fc1 x2)))) bf16.  The residual add is deferred: it is
    fused with the next block's pre-LayerNorm (ops.ln_res_ln with no first LN), so the GEMM epilogue stays a
    plain coalesced bf16 TMA store.  Returns (y, descriptor of the dropout / drop-path still owed to y)."""
    h = ops.linear(x2, layer.fc1.weight, layer.fc1.bias)
    if layer.ffn_layernorm is not None:
        act_drop = ops.dropout_state(h.device).spec(layer.activation_dropout_p) if layer.training else None
        h = ops.layer_norm(h, layer.ffn_layernorm.weight, layer.ffn_layernorm.bias, layer.ffn_layernorm.eps, gelu=True, drop=act_drop)
    else:
        raise NotImplementedError("scale_fc=False")
    y = ops.linear(h, layer.fc2.weight, layer.fc2.bias)
    return y, _branch_drop(layer, y)


def _enter(x, pending, ln):
    """(x, pending = (FFN output of the previous layer, its dropout descriptor)) -> (x + drop(y), ln(x + drop(y)))."""
    if pending is None:
        return x, ln(x)
    y, drop = pending
    return ops.ln_res_ln(y, x, None, None, ln.weight, ln.bias, ln.eps, drop=drop)


def _junction(layer, attn_out, x_res, ln_attn, ln_next):
    """x_new = x_res + drop(ln_attn(attn_out)); y = ln_next(x_new)."""
    if ln_attn is None:
        raise NotImplementedError("scale_attn=False")
    return ops.ln_res_ln(attn_out, x_res, ln_attn.weight, ln_attn.bias, ln_next.weight, ln_next.bias, ln_next.eps,
                         drop=_branch_drop(layer, attn_out))


def _act_dropout_p(cfg):
    """transformer_layer.py:42-46: activation_dropout, falling back to the legacy relu_dropout."""
    p = float(getattr(cfg, "activation_dropout", 0.0) or 0.0)
    if p == 0:
        p = float(getattr(cfg, "relu_dropout", 0.0) or 0.0)
    return p


class TransformerEncoderLayer(nn.Module):
    def __init__(self, args, drop_path_rate=0.0):
        super().__init__()
        cfg = args
        self.cfg = cfg
        _check(cfg, "encoder")
        self.embed_dim = cfg.encoder.embed_dim
        self.self_attn = MultiheadAttention(self.embed_dim, cfg.encoder.attention_heads, dropout=cfg.attention_dropout,
                                            self_attention=True, scale_factor=cfg.attn_scale_factor, scale_heads=cfg.scale_heads)
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.dropout_p = float(cfg.dropout)
        self.activation_dropout_p = _act_dropout_p(cfg)
        self.normalize_before = True
        self.fc1 = nn.Linear(self.embed_dim, cfg.encoder.ffn_embed_dim)
        self.fc2 = nn.Linear(cfg.encoder.ffn_embed_dim, self.embed_dim)
        self.attn_ln = LayerNorm(self.embed_dim) if cfg.scale_attn else None
        self.nh = self.self_attn.num_heads
        self.head_dim = self.self_attn.head_dim
        self.ffn_layernorm = LayerNorm(cfg.encoder.ffn_embed_dim) if cfg.scale_fc else None
        self.w_resid = None
        self.final_layer_norm = LayerNorm(self.embed_dim)
        self.drop_path_rate = float(drop_path_rate)

    def forward(self, x, encoder_padding_mask=None, attn_mask=None, self_attn_bias=None, need_attn=False, modal_mask=None,
                batch_first=False, pending=None, defer=False):
        """x: T x B x C (reference layout) or B x T x C with batch_first=True; fp32 residual stream.
        Internal fast path (used by TransformerEncoder): `pending` = (previous layer's FFN output whose residual add
        is still owed, its dropout descriptor); with defer=True this layer returns ((x, pending'), None) instead of
        adding its own."""
        if not batch_first:
            x = x.transpose(0, 1).contiguous()
        x = ops.to_f32(x)
        x, x1 = _enter(x, pending, self.self_attn_layer_norm)
        a, _ = self.self_attn(x1, key_padding_mask=encoder_padding_mask, attn_bias=self_attn_bias, batch_first=True, causal=attn_mask is not None)
        x, x2 = _junction(self, a, x, self.attn_ln, self.final_layer_norm)
        y, ydrop = _ffn(self, x2)
        if defer:
            return (x, (y, ydrop)), None
        x = ops.add_residual(x, ops.dropout(y, ydrop))
        if not batch_first:
            x = x.transpose(0, 1)
        return x, None


def _forward_incremental_layer(layer, x, encoder_out, encoder_padding_mask, incremental_state, self_kpm, self_bias, cross_bias, batch_first):
    """One decoding step (transformer_layer.py:351-495 with incremental_state): x holds the new position(s) only; the
    self-attention keys / values of earlier positions and the projected encoder output live in `incremental_state`."""
    if not batch_first:
        x = x.transpose(0, 1).contiguous()
        encoder_out = encoder_out.transpose(0, 1).contiguous()
    x = ops.to_f32(x)
    x1 = layer.self_attn_layer_norm(x)
    a, _ = layer.self_attn(x1, key_padding_mask=self_kpm, incremental_state=incremental_state,
                           attn_bias=self_bias if self_bias is not None else False, batch_first=True)
    x, x2 = _junction(layer, a, x, layer.self_attn_ln, layer.encoder_attn_layer_norm)
    c, _ = layer.encoder_attn(x2, key=encoder_out, value=encoder_out, key_padding_mask=encoder_padding_mask,
                              incremental_state=incremental_state, static_kv=True, attn_bias=cross_bias, batch_first=True)
    x, x3 = _junction(layer, c, x, layer.cross_attn_ln, layer.final_layer_norm)
    y, ydrop = _ffn(layer, x3)
    x = ops.add_residual(x, ops.dropout(y, ydrop))
    if not batch_first:
        x = x.transpose(0, 1)
    return x, None, None


class TransformerDecoderLayer(nn.Module):
    def __init__(self, args, no_encoder_attn=False, add_bias_kv=False, add_zero_attn=False, drop_path_rate=0.0):
        super().__init__()
        cfg = args
        self.cfg = cfg
        _check(cfg, "decoder")
        assert not no_encoder_attn and not cfg.cross_self_attention
        self.embed_dim = cfg.decoder.embed_dim
        self.dropout_p = float(cfg.dropout)
        self.activation_dropout_p = _act_dropout_p(cfg)
        self.self_attn = MultiheadAttention(self.embed_dim, cfg.decoder.attention_heads, dropout=cfg.attention_dropout,
                                            self_attention=True, scale_factor=cfg.attn_scale_factor, scale_heads=cfg.scale_heads)
        self.self_attn_ln = LayerNorm(self.embed_dim) if cfg.scale_attn else None
        self.cross_attn_ln = LayerNorm(self.embed_dim) if cfg.scale_attn else None
        self.nh = self.self_attn.num_heads
        self.head_dim = self.self_attn.head_dim
        self.normalize_before = True
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.encoder_attn = MultiheadAttention(self.embed_dim, cfg.decoder.attention_heads, kdim=cfg.encoder.embed_dim,
                                               vdim=cfg.encoder.embed_dim, dropout=cfg.attention_dropout,
                                               encoder_decoder_attention=True, scale_factor=cfg.attn_scale_factor,
                                               scale_heads=cfg.scale_heads)
        self.encoder_attn_layer_norm = LayerNorm(self.embed_dim)
        self.ffn_layernorm = LayerNorm(cfg.decoder.ffn_embed_dim) if cfg.scale_fc else None
        self.w_resid = None
        self.fc1 = nn.Linear(self.embed_dim, cfg.decoder.ffn_embed_dim)
        self.fc2 = nn.Linear(cfg.decoder.ffn_embed_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)
        self.need_attn = True
        self.drop_path_rate = float(drop_path_rate)

    _forward_incremental = _forward_incremental_layer

    def forward(self, x, encoder_out=None, encoder_padding_mask=None, incremental_state=None, prev_self_attn_state=None,
                prev_attn_state=None, self_attn_mask=None, self_attn_padding_mask=None, need_attn=False, need_head_weights=False,
                self_attn_bias=None, cross_attn_bias=None, modal_mask=None, batch_first=False, pending=None, defer=False):
        if prev_self_attn_state is not None or prev_attn_state is not None:
            raise NotImplementedError("prev_*_state injection (ONNX export path) is not supported; use incremental_state")
        if incremental_state is not None:
            return self._forward_incremental(x, encoder_out, encoder_padding_mask, incremental_state, self_attn_padding_mask,
                                             self_attn_bias, cross_attn_bias, batch_first)
        if not batch_first:
            x = x.transpose(0, 1).contiguous()
            encoder_out = encoder_out.transpose(0, 1).contiguous()
        x = ops.to_f32(x)
        x, x1 = _enter(x, pending, self.self_attn_layer_norm)
        # the decoder passes False (not None) when biases are off -> manual path (transformer.py:476-477)
        a, _ = self.self_attn(x1, key_padding_mask=self_attn_padding_mask, attn_bias=self_attn_bias if self_attn_bias is not None else False,
                              batch_first=True, causal=self_attn_mask is not None)
        x, x2 = _junction(self, a, x, self.self_attn_ln, self.encoder_attn_layer_norm)
        c, _ = self.encoder_attn(x2, key=encoder_out, value=encoder_out, key_padding_mask=encoder_padding_mask, static_kv=True,
                                 attn_bias=cross_attn_bias, batch_first=True, causal=False)
        x, x3 = _junction(self, c, x, self.cross_attn_ln, self.final_layer_norm)
        y, ydrop = _ffn(self, x3)
        if defer:
            return (x, (y, ydrop)), None, None
        x = ops.add_residual(x, ops.dropout(y, ydrop))
        if not batch_first:
            x = x.transpose(0, 1)
        return x, None, None
