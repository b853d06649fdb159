"""GeneralistModel / GeneralistModelConfig / OFAEncoderDecoderExecutor with the reference's API
(ofasys/model/ofa.py:41-650): same constructor + initialize(global_dict) two-step, same forward
signature and return values, same parameter names, registered under
("ofasys.model", "unify")."""
import logging
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..adaptor.general import OFAAdaptorConfig
from ..configure import register_config
from ..module import TransformerConfig, init_bert_params
from ..preprocessor import Slot
from .transformer import TransformerDecoder, TransformerEncoder

logger = logging.getLogger(__name__)


@dataclass
class GeneralistModelConfig(TransformerConfig):
    arch: str = "base"
    encode_drop_path_rate: float = 0.0
    decode_drop_path_rate: float = 0.0
    attn_scale_factor: float = 2
    freeze_encoder: bool = False
    freeze_encoder_embedding: bool = False
    freeze_decoder_embedding: bool = False
    add_type_embedding: bool = True
    entangle_position_embedding: bool = False
    sync_bn: bool = False
    scale_attn: bool = True
    scale_fc: bool = True
    scale_heads: bool = True
    scale_resids: bool = False
    use_fused: bool = False
    use_self_attn_bias: bool = True
    adaptor: OFAAdaptorConfig = field(default_factory=OFAAdaptorConfig)
    share_attn_bias: bool = False
    modal_ffn: bool = False

    @classmethod
    def default(cls):
        """The values of ofasys/config/default_model.yaml (what GeneralistModel(cfg=None) loads)."""
        c = cls()
        c.arch = "tiny"
        c.encoder.normalize_before = c.decoder.normalize_before = True
        c.encoder.learned_pos = c.decoder.learned_pos = True
        c.max_source_positions = c.max_target_positions = 1024
        c.share_decoder_input_output_embed = c.share_all_embeddings = True
        c.no_scale_embedding = c.layernorm_embedding = True
        c.activation_fn = "gelu"
        c.dropout = 0.1
        c.attention_dropout = 0.0
        return c


def _arch(cfg, d, layers_enc, layers_dec, heads, resnet):
    cfg.encoder.embed_dim = cfg.decoder.embed_dim = d
    cfg.encoder.ffn_embed_dim = cfg.decoder.ffn_embed_dim = 4 * d
    cfg.decoder.input_dim = cfg.decoder.output_dim = d
    cfg.encoder.layers, cfg.decoder.layers = layers_enc, layers_dec
    cfg.encoder.attention_heads = cfg.decoder.attention_heads = heads
    if hasattr(cfg.adaptor, "image_resnet"):
        cfg.adaptor.image_resnet.resnet_type = resnet


# arch presets (ofa.py:557-650)
def ofa_arch_tiny(cfg): _arch(cfg, 256, 4, 4, 4, "resnet50")
def ofa_arch_medium(cfg): _arch(cfg, 512, 4, 4, 8, "resnet101")
def ofa_arch_base(cfg): _arch(cfg, 768, 6, 6, 12, "resnet101")
def ofa_arch_asr_base(cfg): _arch(cfg, 768, 12, 6, 12, "resnet101")
def ofa_arch_large(cfg): _arch(cfg, 1024, 12, 12, 16, "resnet152")
def ofa_arch_huge(cfg): _arch(cfg, 1280, 24, 12, 16, "resnet152")


class OFAEncoderDecoderExecutor:
    def __init__(self, encoder_name="transformer_encoder", decoder_name="transformer_decoder"):
        self.encoder_name, self.decoder_name = encoder_name, decoder_name

    def forward(self, ofa_model, slots: List[Slot], features_only=False, full_context_alignment=False, alignment_layer=None,
                alignment_heads=None, return_all_hiddens=False, return_encoder_out=False, return_hf_dict=False,
                return_all_attention_weights=False):
        encoder = ofa_model.get_model_by_name(self.encoder_name)
        decoder = ofa_model.get_model_by_name(self.decoder_name)
        encoder_out = encoder([s for s in slots if s.is_src], return_all_hiddens=return_all_hiddens,
                              return_all_attention_weights=return_all_attention_weights)
        decoder_out, extra = decoder([s for s in slots if not s.is_src], encoder_out=encoder_out, features_only=features_only,
                                     full_context_alignment=full_context_alignment, alignment_layer=alignment_layer,
                                     alignment_heads=alignment_heads, return_all_hiddens=return_all_hiddens,
                                     return_all_attention_weights=return_all_attention_weights)
        if return_hf_dict:
            ret = {"last_hidden_state": extra["last_hidden_state"]}
            if return_all_hiddens:
                ret["decoder_hidden_states"] = extra["inner_states"]
            if not features_only:
                ret["decoder_adaptor_out"] = decoder_out
            if return_encoder_out:
                ret["encoder_last_hidden_state"] = encoder_out["encoder_out"]
                if return_all_hiddens:
                    ret["encoder_hidden_states"] = encoder_out["encoder_states"]
            return ret
        if return_encoder_out:
            return decoder_out, extra, encoder_out
        return decoder_out, extra

    def get_logits_from_net_output(self, net_output):
        return net_output["decoder_adaptor_out"] if isinstance(net_output, dict) else net_output[0]

    def get_normalized_probs(self, ofa_model, net_output, log_probs: bool, sample=None):
        logits = self.get_logits_from_net_output(net_output).float()
        return torch.log_softmax(logits, dim=-1) if log_probs else torch.softmax(logits, dim=-1)


@register_config("ofasys.model", "unify", dataclass=GeneralistModelConfig)
class GeneralistModel(nn.Module):
    def __init__(self, cfg: GeneralistModelConfig = None):
        super().__init__()
        if cfg is None:
            cfg = GeneralistModelConfig.default()
        self.cfg = cfg
        if cfg.arch:
            globals()["ofa_arch_" + cfg.arch](cfg)

    def _check_kernel_limits(self):
        """Reject geometries the kernels do not cover at construction time (not at the first forward)."""
        d, f = self.cfg.encoder.embed_dim, self.cfg.encoder.ffn_embed_dim
        if d % self.cfg.encoder.attention_heads != 0 or d // self.cfg.encoder.attention_heads != 64:
            raise NotImplementedError(f"ofasys_b200: head_dim must be 64 (embed_dim {d}, heads {self.cfg.encoder.attention_heads})")
        if d > 1024 or f > 4096:
            raise NotImplementedError(f"ofasys_b200: embed_dim {d} / ffn_embed_dim {f} exceed the LayerNorm-family kernels' limits "
                                      "(embed_dim <= 1024, ffn <= 4096: tiny .. large presets; `huge` (1280 / 5120) is not covered)")

    def initialize(self, global_dict):
        self._check_kernel_limits()
        self.encoder = TransformerEncoder(self.cfg, global_dict)
        self.decoder = TransformerDecoder(self.cfg, global_dict, self.cfg.no_cross_attention)
        self.extra_models = nn.ModuleDict()
        self.active_executor = OFAEncoderDecoderExecutor()
        self.apply(init_bert_params)
        if self.cfg.freeze_encoder:
            self.encoder.requires_grad_(False)
        self.global_dict = global_dict

    def pack_parameters(self):
        """Move the q|k|v (self-attention) and k|v (cross-attention) projection parameters of every attention module into
        shared storages now (otherwise the first forward on the device does it): call after .to(device / dtype) and before
        anything that records parameter addresses (GradArena, FusedAdam tables)."""
        from ..module.multihead_attention import MultiheadAttention

        for mod in self.modules():
            if isinstance(mod, MultiheadAttention) and mod.q_proj.weight.is_cuda:
                names = ("q_proj", "k_proj", "v_proj") if mod.self_attention else ("k_proj", "v_proj")
                mod._cat(names, "weight")
                mod._cat(names, "bias")
        return self

    def get_active_executor(self):
        return self.active_executor

    def set_active_executor(self, executor):
        self.active_executor = executor

    def get_model_by_name(self, model_name: str):
        if model_name == "transformer_encoder":
            return self.encoder
        if model_name == "transformer_decoder":
            return self.decoder
        return self.extra_models[model_name]

    def forward(self, slots: List[Slot], features_only=False, full_context_alignment=False, alignment_layer=None,
                alignment_heads=None, return_all_hiddens=False, return_encoder_out=False, return_hf_dict=False,
                return_all_attention_weights=False):
        return self.active_executor.forward(
            self, slots=slots, features_only=features_only, full_context_alignment=full_context_alignment,
            alignment_layer=alignment_layer, alignment_heads=alignment_heads, return_all_hiddens=return_all_hiddens,
            return_encoder_out=return_encoder_out, return_hf_dict=return_hf_dict,
            return_all_attention_weights=return_all_attention_weights)

    def forward_loss(self, slots: List[Slot], target: torch.Tensor, ignore_index: Optional[int] = None, label_smoothing: float = 0.0,
                     ce_chunk_rows: Optional[int] = None):
        """fwd of the measured path in one call: model + sum-CE criterion (cross_entropy.py:50-67) with the
        tied output projection fused into the loss, so the [B, T, V] logits are never kept in fp32.
        label_smoothing > 0: the label-smoothed criterion (label_smoothed_cross_entropy.py:62-92,175-191).
        ce_chunk_rows: run projection + criterion that many target rows at a time (no [B * T, V] scratch at all; see
        ops.linear_cross_entropy).  Returns (loss_sum, sample_size = ntokens is left to the caller)."""
        feats, _ = self.forward(slots, features_only=True)
        pad = self.global_dict.pad() if ignore_index is None else ignore_index
        return ops.linear_cross_entropy(feats, self.decoder.adaptor.embed_tokens.weight, target, pad, label_smoothing, chunk_rows=ce_chunk_rows)

    def forward_backward_split(self, slots: List[Slot], target: torch.Tensor, ignore_index: Optional[int] = None):
        """fwd + bwd of the measured path with the backward cut at the encoder/decoder boundary, for data-parallel
        overlap (the reference gets this from c10d DDP's bucket hooks, distributed_model_dispatcher.py:49-75):

            loss, early, finish = model.forward_backward_split(slots, target)
            ... start averaging `early` (decoder-side parameters; their .grad is final) ...
            finish()        # encoder backward; afterwards every other parameter's .grad is final

        Parameters shared with the encoder (the tied embedding) are finished by `finish()`."""
        enc_out = self.encoder([s for s in slots if s.is_src])
        # cut the autograd graph at the boundary: the decoder sees detached leaves, stage 2 feeds their gradients back
        cut, leaves, dec_in = [], [], {}

        def leaf(t):
            if not (torch.is_tensor(t) and t.requires_grad):
                return t
            d = t.detach().requires_grad_(True)
            cut.append(t)
            leaves.append(d)
            return d

        for k, v in enc_out.items():
            if k == "encoder_out" and "_encoder_out_bt" in enc_out:
                continue  # T x B x C view of `_encoder_out_bt`, rebuilt below from the detached leaf
            dec_in[k] = [leaf(t) for t in v] if isinstance(v, (list, tuple)) else leaf(v)
        if "_encoder_out_bt" in enc_out:
            dec_in["encoder_out"] = [dec_in["_encoder_out_bt"].transpose(0, 1)]
        feats, _ = self.decoder([s for s in slots if not s.is_src], encoder_out=dec_in, features_only=True)
        pad = self.global_dict.pad() if ignore_index is None else ignore_index
        loss = ops.linear_cross_entropy(feats, self.decoder.adaptor.embed_tokens.weight, target, pad)
        enc_ids = {id(p) for p in self.encoder.parameters()}
        early = [p for p in self.parameters() if p.requires_grad and id(p) not in enc_ids]
        shared = [p for p in self.decoder.parameters() if p.requires_grad and id(p) in enc_ids]
        grads = torch.autograd.grad(loss, leaves + early + shared, allow_unused=True)
        for p, g in zip(early + shared, grads[len(cut):]):
            p.grad = g
        todo = [(t, g) for t, g in zip(cut, grads[:len(cut)]) if g is not None]

        def finish():
            if todo:
                torch.autograd.backward([t for t, _ in todo], [g for _, g in todo])

        return loss, early, finish

    def get_normalized_probs(self, net_output, log_probs: bool, sample=None):
        return self.active_executor.get_normalized_probs(self, net_output, log_probs, sample)

    def get_targets(self, sample, net_output):
        return sample["target"]

    def update_sample(self, sample):
        sample = self.encoder.adaptor.update_sample(sample)
        return self.decoder.adaptor.update_sample(sample)

    def set_num_updates(self, num_updates):
        pass

    def max_positions(self):
        return (self.encoder.max_positions(), self.decoder.max_positions())

    def max_decoder_positions(self):
        return self.decoder.max_positions()
