"""TransformerEncoder / TransformerDecoder (ofasys/model/transformer.py:33-539) over the fused layers.

Internal layout is batch-major [B, T, C] (fp32 residual stream, bf16 GEMM operands); the dict the
encoder returns keeps the reference's keys and T x B x C views for `encoder_out`.
"""
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..adaptor import AdaptorOutput, OFAGeneralAdaptor
from ..module import LayerNorm, TransformerDecoderLayer, TransformerEncoderLayer
from ..preprocessor import Slot


class TransformerEncoder(nn.Module):
    def __init__(self, cfg, dictionary):
        super().__init__()
        self.cfg = cfg
        self.dictionary = dictionary
        self.register_buffer("version", torch.Tensor([3]))
        OFAGeneralAdaptor._embed_tokens = None  # a new model builds a new shared embedding (transformer.py:48)
        self.adaptor = OFAGeneralAdaptor(cfg, dictionary, True)
        assert cfg.encoder.layerdrop == 0.0, "LayerDrop is not on the hot path"
        dpr = torch.linspace(0, cfg.encode_drop_path_rate, cfg.encoder.layers)
        self.layers = nn.ModuleList([TransformerEncoderLayer(cfg, drop_path_rate=float(dpr[i])) for i in range(cfg.encoder.layers)])
        self.layer_norm = LayerNorm(cfg.encoder.embed_dim) if cfg.encoder.normalize_before else None
        self._uses_dropout = any(float(getattr(cfg, k, 0.0) or 0.0) > 0 for k in
                                 ("dropout", "attention_dropout", "activation_dropout", "relu_dropout", "encode_drop_path_rate"))

    def forward(self, slots: List[Slot], return_all_hiddens: bool = False, return_all_attention_weights: bool = False):
        if len(slots) == 0:
            return None
        if return_all_attention_weights:
            raise NotImplementedError("attention maps are never materialised by the fused kernel")
        if self.training and self._uses_dropout:
            ops.dropout_state().next_step()  # the encoder opens a step: fresh masks, call-site counter restarted
        embed, masks, pos, biases, _ = self.adaptor(slots)
        x = embed  # padded rows already zeroed by the adaptor kernels (transformer.py:109-112)
        states = [x.transpose(0, 1)] if return_all_hiddens else []
        defer = not return_all_hiddens and self.layer_norm is not None  # fast path: residual adds fused into the next LN
        pending = None
        for idx, layer in enumerate(self.layers):
            bias = None
            if self.cfg.use_self_attn_bias:
                bias = biases[0] if self.cfg.share_attn_bias else biases[idx]
            # an all-False padding mask is numerically identical to the reference's `None` (no host sync on masks.any())
            out, _ = layer(x, encoder_padding_mask=masks, self_attn_bias=bias, batch_first=True, pending=pending, defer=defer)
            if defer:
                x, pending = out
            else:
                x = out
                states.append(x.transpose(0, 1))
        if pending is not None:
            x, xb = ops.ln_res_ln(pending[0], x, None, None, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps, drop=pending[1])
        else:
            xb = self.layer_norm(x) if self.layer_norm is not None else ops.to_bf16(x)  # bf16 [B, S, C]
        return {
            "encoder_out": [xb.transpose(0, 1)],  # T x B x C view
            "encoder_padding_mask": [masks],  # B x T
            "encoder_embedding": [embed],  # B x T x C
            "encoder_states": states,
            "position_embeddings": [pos],  # B x T x C
            "encoder_attention_weights": [],
            "_encoder_out_bt": xb,  # B x T x C, contiguous (what the decoder layers consume)
        }

    def reorder_encoder_out(self, encoder_out: Dict[str, List[torch.Tensor]], new_order):
        def sel(lst, dim):
            return [t.index_select(dim, new_order) for t in lst if t is not None]

        out = {
            "encoder_out": sel(encoder_out["encoder_out"], 1),
            "encoder_padding_mask": sel(encoder_out["encoder_padding_mask"], 0),
            "encoder_embedding": sel(encoder_out["encoder_embedding"], 0),
            "encoder_states": sel(encoder_out["encoder_states"], 1),
            "position_embeddings": sel(encoder_out["position_embeddings"], 0),
            "encoder_attention_weights": [],
        }
        out["_encoder_out_bt"] = encoder_out["_encoder_out_bt"].index_select(0, new_order)
        return out

    def max_positions(self):
        return self.cfg.max_source_positions


class TransformerDecoder(nn.Module):
    def __init__(self, cfg, dictionary, no_encoder_attn=False):
        super().__init__()
        self.cfg = cfg
        self.dictionary = dictionary
        self.register_buffer("version", torch.Tensor([3]))
        self.adaptor = OFAGeneralAdaptor(cfg, dictionary, False)
        self.share_input_output_embed = cfg.share_decoder_input_output_embed
        self.num_attention_heads = cfg.decoder.attention_heads
        embed_dim = cfg.decoder.embed_dim
        self.embed_dim = embed_dim
        self.output_embed_dim = int(cfg.decoder.output_dim)
        assert self.output_embed_dim == embed_dim, "project_out_dim is not on the OFA path"
        if cfg.use_self_attn_bias:
            self.cross_pos_q_linear = nn.Linear(embed_dim, embed_dim)
            self.cross_pos_k_linear = nn.Linear(embed_dim, embed_dim)
        assert cfg.decoder.layerdrop == 0.0
        # quirk 4: the decoder's drop-path schedule uses the *encoder* rate and layer count
        dpr = torch.linspace(0, cfg.encode_drop_path_rate, cfg.encoder.layers)
        self.layers = nn.ModuleList(
            [TransformerDecoderLayer(cfg, no_encoder_attn, drop_path_rate=float(dpr[i])) for i in range(cfg.decoder.layers)]
        )
        self.num_layers = len(self.layers)
        self.layer_norm = LayerNorm(embed_dim) if cfg.decoder.normalize_before else None
        self.project_out_dim = None
        self.adaptive_softmax = None

    def get_cross_pos_info(self, tgt_pos_embed, src_pos_embed, dense=True):
        """abs position term of the cross-attention bias (transformer.py:280-299): one dense batch-invariant [H, T, S] term
        shared by all decoder layers (dense=False: as extra QK columns of the mma.sync kernels, incremental decoding)."""
        pq = ops.linear(ops.to_bf16(tgt_pos_embed[:1]), self.cross_pos_q_linear.weight, self.cross_pos_q_linear.bias)
        pk = ops.linear(ops.to_bf16(src_pos_embed[:1]), self.cross_pos_k_linear.weight, self.cross_pos_k_linear.bias)
        return ops.PositionBias(pq, pk, abs=ops.abs_pos(pq, pk, self.num_attention_heads) if dense else None)

    def forward(self, slots: List[Slot], encoder_out: Optional[Dict[str, List[torch.Tensor]]] = None, incremental_state=None,
                features_only: bool = False, full_context_alignment: bool = False, alignment_layer: Optional[int] = None,
                alignment_heads: Optional[int] = None, return_all_hiddens: bool = False, return_all_attention_weights: bool = False):
        x, extra = self.extract_features(slots, encoder_out=encoder_out, incremental_state=incremental_state,
                                         full_context_alignment=full_context_alignment, return_all_hiddens=return_all_hiddens,
                                         return_all_attention_weights=return_all_attention_weights)
        extra["last_hidden_state"] = x
        if not features_only:
            return self.adaptor.forward_output(x, extra, slots)
        return x, extra

    def extract_features(self, slots, encoder_out, incremental_state=None, full_context_alignment=False, alignment_layer=None,
                         alignment_heads=None, return_all_hiddens=False, return_all_attention_weights=False):
        if return_all_attention_weights:
            raise NotImplementedError("attention maps are never materialised by the fused kernel")
        if incremental_state is not None:
            return self._extract_features_incremental(slots, encoder_out, incremental_state)
        embed, masks, pos, biases, _ = self.adaptor(slots)
        enc = encoder_out["_encoder_out_bt"] if "_encoder_out_bt" in encoder_out else encoder_out["encoder_out"][0].transpose(0, 1).contiguous()
        enc_mask = encoder_out["encoder_padding_mask"][0] if encoder_out["encoder_padding_mask"] else None
        cross_bias = None
        if not self.cfg.entangle_position_embedding:
            cross_bias = self.get_cross_pos_info(pos, encoder_out["position_embeddings"][0])
        x = embed
        inner = [x.transpose(0, 1)] if return_all_hiddens else []
        causal = not full_context_alignment
        defer = not return_all_hiddens and self.layer_norm is not None
        pending = None
        for idx, layer in enumerate(self.layers):
            bias = None
            if self.cfg.use_self_attn_bias:
                bias = biases[0] if self.cfg.share_attn_bias else biases[idx]
            out, _, _ = layer(x, enc, enc_mask, self_attn_mask=True if causal else None, self_attn_padding_mask=masks,
                              self_attn_bias=bias, cross_attn_bias=cross_bias, batch_first=True, pending=pending, defer=defer)
            if defer:
                x, pending = out
            else:
                x = out
                inner.append(x.transpose(0, 1))
        if pending is not None:
            _, x = ops.ln_res_ln(pending[0], x, None, None, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps, drop=pending[1])
        else:
            x = self.layer_norm(x) if self.layer_norm is not None else ops.to_bf16(x)  # bf16 B x T x C
        return x, {"attn": [None], "inner_states": inner, "decoder_attentions": [], "cross_attentions": []}

    def _extract_features_incremental(self, slots, encoder_out, incremental_state):
        """One generation step (transformer.py:417-522 with incremental_state).  As in the reference the
        adaptor sees the whole prefix (embeddings / position terms of every position so far) and the last position is
        sliced out (:444-447, :474-475); the layers then run on that single row against the caches in
        `incremental_state` (a plain dict owned by the caller, e.g. the sequence generator)."""
        embed, masks, pos, biases, _ = self.adaptor(slots)
        enc = encoder_out["_encoder_out_bt"] if "_encoder_out_bt" in encoder_out else encoder_out["encoder_out"][0].transpose(0, 1).contiguous()
        enc_mask = encoder_out["encoder_padding_mask"][0] if encoder_out["encoder_padding_mask"] else None
        cross_bias = None
        if not self.cfg.entangle_position_embedding:
            cb = self.get_cross_pos_info(pos, encoder_out["position_embeddings"][0], dense=False)
            cross_bias = ops.PositionBias(cb.pq[:, -1:].contiguous(), cb.pk)
        x = embed[:, -1:].contiguous()
        last_mask = masks[:, -1:]
        step_biases = None
        if self.cfg.use_self_attn_bias:
            step_biases = []
            for b in biases:
                idx = None if b.rp_idx is None else b.rp_idx[-1:, :].contiguous()
                step_biases.append(ops.PositionBias(b.pq[:, -1:].contiguous(), b.pk, idx, b.table))
        for idx, layer in enumerate(self.layers):
            bias = None
            if step_biases is not None:
                bias = step_biases[0] if self.cfg.share_attn_bias else step_biases[idx]
            x, _, _ = layer(x, enc, enc_mask, incremental_state=incremental_state, self_attn_padding_mask=last_mask,
                            self_attn_bias=bias, cross_attn_bias=cross_bias, batch_first=True)
        x = self.layer_norm(x) if self.layer_norm is not None else ops.to_bf16(x)
        return x, {"attn": [None], "inner_states": [], "decoder_attentions": [], "cross_attentions": []}

    def reorder_incremental_state(self, incremental_state, new_order):
        """Beam search: permute every attention cache along the batch (sequence_generator.py reorder_incremental_state;
        multihead_attention.py:393-409)."""
        for layer in self.layers:
            layer.self_attn.reorder_incremental_state(incremental_state, new_order)
            layer.encoder_attn.reorder_incremental_state(incremental_state, new_order)
        return incremental_state

    reorder_incremental_state_scripting = reorder_incremental_state

    def max_positions(self):
        return self.cfg.max_target_positions
